set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2d_pytest.log; cat $O/r2d_pytest.log
bash tests/tools/sanitize.sh $O/sanitize 2>&1 | tail -20
python tests/tools/bench_configs.py 2>&1 | grep -E "sweep_T25|config3"
