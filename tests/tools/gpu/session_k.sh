set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${NGPU:-8}
nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().split('\n')[-1])
print('$2', 'ms %.3f p50 %.3f p99 %.3f'%(d['ms_per_step'],d['p50_ms'],d['p99_ms']), [round(x,3) for x in d['run']['ms_per_step_by_rank']]); print('   ', d['run']['ms_by_step_rank0']); print('   clocks', d['clocks'])
"; }
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu > $O/r2k_a.json 2> $O/r2k_a.err; show $O/r2k_a.json base
BENCH_SAMPLER_PERIOD=0.1 timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu > $O/r2k_b.json 2> $O/r2k_b.err; show $O/r2k_b.json sampler100ms
BENCH_SAMPLER_PERIOD=0.1 BENCH_PIN=1 timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu > $O/r2k_c.json 2> $O/r2k_c.err; show $O/r2k_c.json pin
BENCH_SAMPLER_PERIOD=0.1 timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --gather nccl > $O/r2k_d.json 2> $O/r2k_d.err; show $O/r2k_d.json nccl
