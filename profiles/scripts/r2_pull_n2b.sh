#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NMGPU_PULL_TRACE=1 timeout 900 $T --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 2 --ct-vmult owned-pull --no-e2e --no-check > $O/r2ad_bench_n2_pull.json 2> $O/r2ad_bench_n2_pull.err; echo "pull rc $?"
grep -a "pull trace" $O/r2ad_bench_n2_pull.err | tail -3
