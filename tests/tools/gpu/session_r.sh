set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
for c in 100 86 75; do
  echo "carveout $c" | tee -a $O/r2r_ab.log
  JMPC_DEBUG=1 JMPC_CARVEOUT=$c JMPC_LIB=$PWD/build/variants/lib_exp.so python tests/tools/ab_bench.py 2>&1 | grep -E "ms |geometry" | sort -u | tee -a $O/r2r_ab.log
done
