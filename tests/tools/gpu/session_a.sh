set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
python tests/tools/bench_configs.py > gpurun_out/r2a_configs.log 2>&1; tail -8 gpurun_out/r2a_configs.log
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
