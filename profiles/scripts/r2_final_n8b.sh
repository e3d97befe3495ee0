#!/bin/bash
# round 2, last 8-GPU pass: the default line only (config 5, checks, e2e with the cube input format)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 1200 $T --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 2 > $O/r2yy_bench_n8.json 2> $O/r2yy_bench_n8.err; echo "n8 rc $?"
tail -c 300 $O/r2yy_bench_n8.err
