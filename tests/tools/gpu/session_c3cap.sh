set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
c=3
python tests/tools/profile_step.py $c > $O/r2f_plain_$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mpc_step_kernel -s 3 -c 1 -f -o $O/r2_step_c$c python tests/tools/profile_step.py $c > $O/r2f_ncu_$c.log 2>&1
tail -1 $O/r2f_ncu_$c.log
ncu -i $O/r2_step_c$c.ncu-rep --page raw --csv > $O/r2_step_c${c}_raw.csv 2>/dev/null
python profiles/tools/ncu_traffic.py $O/r2_step_c3_raw.csv B65536_T13 "gpurun_out/r2_step_c3.ncu-rep (ncu --set full --clock-control none), one launch" profiles/r2_traffic.json
cp profiles/r2_traffic.json $O/r2_traffic.json
python bench.py --config 3 --steps 10 --warmup 3 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; tail -c 200 $O/r2_bench_c3.json
