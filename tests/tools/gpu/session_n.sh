set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r2n_pytest.log; cat $O/r2n_pytest.log
for v in base opt2_rall_ptx opt3 opt3_t8b4 opt3_t8b5 opt3_t8b7 opt3; do
  JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep "ms " | tee -a $O/r2n_ab.log
done
for c in 3 2; do
  python tests/tools/profile_step.py $c > $O/r2n_plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:mpc_step_kernel -s 3 -c 1 -f -o $O/r2n_step_c$c python tests/tools/profile_step.py $c > $O/r2n_ncu_$c.log 2>&1
  tail -2 $O/r2n_ncu_$c.log
done
ls -la $O/*.ncu-rep
