#!/bin/bash
# round 2, after the late merge / transpose / SpMV changes: 2 GPUs -- the multi-GPU test and the default bench line with its checks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -k two_gpus -q > $O/r2x_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2x_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err; echo "n2 rc $?"
tail -c 400 $O/r2x_bench_n2.err
