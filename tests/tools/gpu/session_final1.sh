# Final single-GPU pass of round 2: tests, smoke, ncu captures (traffic json refreshed before the benches read it),
# launch lists, bench lines of configs 2-5 and the reference arm, planner / episode throughput.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_gputest.log; cat $O/r2_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2_smoke.log
for c in 2 3; do
  python tests/tools/profile_step.py $c > $O/r2f_plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:mpc_step_kernel -s 3 -c 1 -f -o $O/r2_step_c$c python tests/tools/profile_step.py $c > $O/r2f_ncu_$c.log 2>&1
  tail -1 $O/r2f_ncu_$c.log
  ncu -i $O/r2_step_c$c.ncu-rep --page raw --csv > $O/r2_step_c${c}_raw.csv 2>/dev/null
done
python profiles/tools/ncu_traffic.py $O/r2_step_c2_raw.csv B4096_T20 "gpurun_out/r2_step_c2.ncu-rep (ncu --set full --clock-control none), one launch" profiles/r2_traffic.json
python profiles/tools/ncu_traffic.py $O/r2_step_c3_raw.csv B65536_T13 "gpurun_out/r2_step_c3.ncu-rep (ncu --set full --clock-control none), one launch" profiles/r2_traffic.json
cp profiles/r2_traffic.json $O/r2_traffic.json
python tests/tools/profile_all_kernels.py > $O/r2f_allk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none --csv --log-file $O/r2_all_kernels.csv python tests/tools/profile_all_kernels.py > $O/r2f_allk_ncu.log 2>&1; wc -l $O/r2_all_kernels.csv
python bench.py --steps 30 --warmup 5 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 400 $O/r2_bench_c2.json
python bench.py --config 3 --steps 10 --warmup 3 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; tail -c 300 $O/r2_bench_c3.json
python bench.py --config 4 --steps 5 --warmup 3 > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err; tail -c 300 $O/r2_bench_c4.json
python bench.py --config 5 --steps 5 --warmup 3 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err; tail -c 300 $O/r2_bench_c5.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_c2.json 2> $O/r2_bench_ref_c2.err; cut -c1-300 $O/r2_bench_ref_c2.json
python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > $O/r2f_benchncu.log 2>&1; wc -l $O/r2_bench_launches.csv
python tests/tools/bench_planner.py 2>&1 | tail -4 | tee $O/r2_planner.log
python tests/tools/bench_episodes.py 4096 13 2>&1 | tail -1 | tee $O/r2_episodes.log; python tests/tools/bench_episodes.py 65536 13 2>&1 | tail -1 | tee -a $O/r2_episodes.log
python tests/tools/bench_configs.py 2>&1 | grep -E "^(config|sweep)" | tee $O/r2_configs.log
