set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py --config 4 --steps 5 --warmup 3 > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err; tail -c 150 $O/r2_bench_c4.json
python bench.py --config 5 --steps 5 --warmup 3 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err; tail -c 150 $O/r2_bench_c5.json
