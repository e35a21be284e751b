run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stitch_only_ms_max_over_ranks'])"; }
run "X=1"
run "NCCL_MAX_NCHANNELS=4"
run "NCCL_MAX_NCHANNELS=2 NCCL_NTHREADS=256"
run "NCCL_MAX_NCHANNELS=8 NCCL_NTHREADS=128"
run "NCCL_PROTO=LL128"
