set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2g_n${N}_c2_every.json 2> $O/r2g_n${N}_c2_every.err; tail -c 400 $O/r2g_n${N}_c2_every.json; tail -3 $O/r2g_n${N}_c2_every.err
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --gather deferred > $O/r2g_n${N}_c2_deferred.json 2> $O/r2g_n${N}_c2_deferred.err; tail -c 400 $O/r2g_n${N}_c2_deferred.json; tail -3 $O/r2g_n${N}_c2_deferred.err
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --gather nccl > $O/r2g_n${N}_c2_nccl.json 2> $O/r2g_n${N}_c2_nccl.err; tail -c 300 $O/r2g_n${N}_c2_nccl.json; tail -3 $O/r2g_n${N}_c2_nccl.err
timeout 600 $TR bench.py --gpus $N --config 5 --steps 3 --warmup 3 > $O/r2g_n${N}_c5.json 2> $O/r2g_n${N}_c5.err; tail -c 400 $O/r2g_n${N}_c5.json; tail -3 $O/r2g_n${N}_c5.err
timeout 300 $TR bench.py --gpus $N --config 3 --steps 5 --warmup 3 > $O/r2g_n${N}_c3.json 2> $O/r2g_n${N}_c3.err; tail -c 300 $O/r2g_n${N}_c3.json; tail -3 $O/r2g_n${N}_c3.err
