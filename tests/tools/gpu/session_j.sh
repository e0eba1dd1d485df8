set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in flags barrier; do
  timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --gather $mode --no-cpu > $O/r2j_n${N}_c2_$mode.json 2> $O/r2j_n${N}_c2_$mode.err; tail -c 200 $O/r2j_n${N}_c2_$mode.json; tail -3 $O/r2j_n${N}_c2_$mode.err
done
timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu > $O/r2j_n${N}_c5.json 2> $O/r2j_n${N}_c5.err; tail -c 200 $O/r2j_n${N}_c5.json; tail -3 $O/r2j_n${N}_c5.err
