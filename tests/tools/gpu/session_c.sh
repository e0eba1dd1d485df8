set -x
cd $GRAFT_REPO_ROOT
for v in base pad5 pad2 mb3 base pad5; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms "
done
(cd build/r1_tree && python tests/tools/ab_bench.py 2>&1 | grep "ms ")
