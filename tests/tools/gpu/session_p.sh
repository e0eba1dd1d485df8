set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
JMPC_LIB=$PWD/build/variants/lib_opt4.so python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/r2p_pytest.log; cat $O/r2p_pytest.log
for v in opt3b opt4 opt3b opt4; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a $O/r2p_ab.log
done
JMPC_LIB=$PWD/build/variants/libjmpc_cycles.so python tests/tools/cycles_probe.py 2>&1 | tee $O/r2p_cycles.log
