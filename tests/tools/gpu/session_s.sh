set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
JMPC_DEBUG=1 python tests/tools/ab_bench.py 2>&1 | grep -E "ms |geometry" | sort -u | tee $O/r2s_ab.log
python tests/tools/ab_bench.py 2>&1 | grep -E "ms " | tee -a $O/r2s_ab.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/r2s_pytest.log
python bench.py --steps 30 --warmup 5 --no-cpu > $O/r2s_bench_c2.json 2> $O/r2s_bench_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_c2.json').read().strip().split('\n')[-1])
print('c2 value %.4g ms %.4g p50 %.4g p99 %.4g e2e %.4g frac %.4f single %.3f hints %.4g'%(d['value'],d['ms_per_step'],d['p50_ms'],d['p99_ms'],d['e2e']['value'],d['roofline']['frac'],d['single_instance_step_ms']['p50'],d['with_history_hints']['ms_per_step']))
PY
