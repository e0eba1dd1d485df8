set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/r2x_pytest.log
for v in final t8b5 t8b7 t8b8 final; do
  JMPC_DEBUG=1 JMPC_LIB=$PWD/build/variants/lib_$v.so python tests/tools/ab_bench.py 2>&1 | grep -E "sweep_T8|T=8 " | sort -u | tee -a $O/r2x_ab.log
done
