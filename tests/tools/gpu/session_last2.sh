set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/r2_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2_smoke.log
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu > $O/r2last_bench_c3.json 2> $O/r2last_bench_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2last_bench_c3.json').read().strip().split('\n')[-1])
print('c3 value %.4g ms %.4g p99 %.4g e2e %.4g frac %.4f'%(d['value'],d['ms_per_step'],d['p99_ms'],d['e2e']['value'],d['roofline']['frac']))
PY
