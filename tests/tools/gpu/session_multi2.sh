# Multi-GPU bench lines, second pass (flags = every step waits for its own table): NGPU=4|8, MODES="flags deferred barrier"
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${NGPU:-4}
MODES=${MODES:-"flags deferred barrier"}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[2], 'value %.4g ms %.3f p50 %.3f p99 %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d['p50_ms'],d['p99_ms'],d['e2e']['value']), [round(x,3) for x in d['run']['ms_per_step_by_rank']])
PY
}
if [ -n "$CALIB" ]; then
python bench.py --steps 40 --warmup 5 --no-cpu > $O/r2_n${N}box_n1_c2.json 2> $O/r2_n${N}box_n1_c2.err; show $O/r2_n${N}box_n1_c2.json "N=1 on this box"
fi
for m in $MODES; do
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --gather $m > $O/r2_n${N}_c2_$m.json 2> $O/r2_n${N}_c2_$m.err; show $O/r2_n${N}_c2_$m.json $m || tail -5 $O/r2_n${N}_c2_$m.err
done
if [ -n "$C5" ]; then
timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu > $O/r2_n${N}_c5.json 2> $O/r2_n${N}_c5.err; show $O/r2_n${N}_c5.json c5 || tail -5 $O/r2_n${N}_c5.err
fi
