# Multi-GPU bench lines of the final build: NGPU=2|4|8 bash tests/tools/gpu/session_multi.sh   (gpurun --gpus N)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[2], 'value %.4g ms %.3f p50 %.3f p99 %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d['p50_ms'],d['p99_ms'],d['e2e']['value']), [round(x,3) for x in d['run']['ms_per_step_by_rank']], d['run'].get('gather'))
PY
}
nvidia-smi -L | head -8
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu > $O/r2_n${N}_c2_flags.json 2> $O/r2_n${N}_c2_flags.err; show $O/r2_n${N}_c2_flags.json flags || tail -5 $O/r2_n${N}_c2_flags.err
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --gather barrier > $O/r2_n${N}_c2_barrier.json 2> $O/r2_n${N}_c2_barrier.err; show $O/r2_n${N}_c2_barrier.json barrier || tail -5 $O/r2_n${N}_c2_barrier.err
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --gather nccl > $O/r2_n${N}_c2_nccl.json 2> $O/r2_n${N}_c2_nccl.err; show $O/r2_n${N}_c2_nccl.json nccl || tail -5 $O/r2_n${N}_c2_nccl.err
timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu > $O/r2_n${N}_c5.json 2> $O/r2_n${N}_c5.err; show $O/r2_n${N}_c5.json c5 || tail -5 $O/r2_n${N}_c5.err
