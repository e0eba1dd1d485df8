set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2b_pytest.log; cat $O/r2b_pytest.log
for i in 1 2; do
  (cd build/r1_tree && python bench.py --steps 30 --warmup 5 --no-cpu) 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('R1 ', d['ms_per_step'], d['p50_ms'], d['with_history_hints']['ms_per_step'], d['single_instance_step_ms'])"
  python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('NEW', d['ms_per_step'], d['p50_ms'], d['with_history_hints']['ms_per_step'], d['single_instance_step_ms'])"
done
python bench.py --steps 20 --warmup 5 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err; tail -c 600 $O/r2b_bench_c2.json
python bench.py --config 3 --steps 10 --warmup 3 > $O/r2b_bench_c3.json 2> $O/r2b_bench_c3.err; tail -c 1200 $O/r2b_bench_c3.json; tail -3 $O/r2b_bench_c3.err
python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > $O/r2b_bench_c4.json 2> $O/r2b_bench_c4.err; tail -c 600 $O/r2b_bench_c4.json; tail -3 $O/r2b_bench_c4.err
python bench.py --config 5 --steps 5 --warmup 3 > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; tail -c 1500 $O/r2b_bench_c5.json; tail -3 $O/r2b_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_ref_c2.json 2> $O/r2b_ref_c2.err; cut -c1-300 $O/r2b_ref_c2.json
