set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python bench.py --steps 30 --warmup 5 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; tail -c 300 $O/r2_bench_c2.json
JMPC_LIB=$PWD/build/variants/libjmpc_cycles.so python tests/tools/cycles_probe.py 2>&1 | tee $O/r2z_cycles.log
python tests/tools/tail_probe.py 2>&1 | tee $O/r2z_tail.log
