set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2e_pytest.log; cat $O/r2e_pytest.log
python tests/tools/bench_planner.py 2>&1 | tail -6
python tests/tools/profile_all_kernels.py > $O/r2e_profile_plain.log 2>&1; tail -3 $O/r2e_profile_plain.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none --csv --log-file $O/r2_all_kernels.csv python tests/tools/profile_all_kernels.py > $O/r2e_ncu_all.log 2>&1; tail -2 $O/r2e_ncu_all.log; wc -l $O/r2_all_kernels.csv
