set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r2_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2_smoke.log
python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2last_bench_c2.json 2> $O/r2last_bench_c2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2last_bench_c2.json').read().strip().split('\n')[-1])
print('c2 value %.4g ms %.4g p99 %.4g e2e %.4g single %.3f'%(d['value'],d['ms_per_step'],d['p99_ms'],d['e2e']['value'],d['single_instance_step_ms']['p50']))
PY
python tests/tools/bench_episodes.py 4096 13 2>&1 | tail -1 | cut -c1-200
