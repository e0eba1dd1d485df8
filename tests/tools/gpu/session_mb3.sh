set -x
cd $GRAFT_REPO_ROOT
for v in final mb3 final mb3; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a gpurun_out/r2_mb3_ab.log
done
