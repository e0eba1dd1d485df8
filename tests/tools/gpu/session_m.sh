set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
JMPC_LIB=$PWD/build/variants/lib_opt2_rsymv_ptx.so python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2m_pytest.log; cat $O/r2m_pytest.log
for v in base opt2 opt2_rsymv opt2_rsymv_ptx opt2_rall_ptx base opt2_rsymv_ptx; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a $O/r2m_ab.log
done
