set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/r2w_pytest.log
for v in mom cum mom cum; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a $O/r2w_ab.log
done
