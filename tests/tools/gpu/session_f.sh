set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tests/tools/bench_planner.py 2>&1 | tail -5
JMPC_LIB=$PWD/build/variants/lib_exp.so JMPC_SCHED_MAX_WAVES=16 python tests/tools/ab_schedule.py 2>&1 | grep "ms "
JMPC_LIB=$PWD/build/variants/lib_exp.so JMPC_SCHED_MAX_WAVES=1000 python tests/tools/ab_schedule.py 2>&1 | grep "ms "
for c in 2 3; do
  python tests/tools/profile_step.py $c > $O/r2f_plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:mpc_step_kernel -s 3 -c 1 -f -o $O/r2_step_c$c python tests/tools/profile_step.py $c > $O/r2f_ncu_$c.log 2>&1
  tail -2 $O/r2f_ncu_$c.log
done
ls -la $O/*.ncu-rep
