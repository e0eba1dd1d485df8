set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2o_pytest.log; cat $O/r2o_pytest.log
JMPC_LIB=$PWD/build/variants/libjmpc_cycles.so python tests/tools/cycles_probe.py 2>&1 | tee $O/r2o_cycles.log
python tests/tools/tail_probe.py 2>&1 | tee $O/r2o_tail.log
python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee $O/r2o_ab.log
