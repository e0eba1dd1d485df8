set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/r2h_pytest.log; cat $O/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > $O/r2h_bench_c2.json 2> $O/r2h_bench_c2.err; tail -c 300 $O/r2h_bench_c2.json
python bench.py --config 3 --steps 10 --warmup 3 > $O/r2h_bench_c3.json 2> $O/r2h_bench_c3.err; tail -c 300 $O/r2h_bench_c3.json
python bench.py --config 4 --steps 5 --warmup 3 > $O/r2h_bench_c4.json 2> $O/r2h_bench_c4.err; tail -c 300 $O/r2h_bench_c4.json
python bench.py --config 5 --steps 5 --warmup 3 > $O/r2h_bench_c5.json 2> $O/r2h_bench_c5.err; tail -c 300 $O/r2h_bench_c5.json
python tests/tools/ab_schedule.py 2>&1 | grep "ms "
python tests/tools/bench_episodes.py 4096 13 2>&1 | tail -1; python tests/tools/bench_episodes.py 65536 13 2>&1 | tail -1
