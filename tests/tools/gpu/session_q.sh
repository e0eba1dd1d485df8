set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
JMPC_LIB=$PWD/build/variants/lib_opt5.so python -m pytest tests/test_gpu_linalg.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -3 > $O/r2q_pytest.log; cat $O/r2q_pytest.log
for v in opt3b opt5 opt3b opt5; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a $O/r2q_ab.log
done
JMPC_LIB=$PWD/build/variants/libjmpc_cycles.so python tests/tools/cycles_probe.py 2>&1 | tee $O/r2q_cycles.log
