#!/bin/bash
# round 2: 8 GPUs -- config 5 (sphere cycle 10, ~1e9 pairs) with checks, and the reference arm's sample at that cycle
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r2y_gpu.txt; free -g >> $O/r2y_gpu.txt; nproc >> $O/r2y_gpu.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
NMGPU_BENCH_TRACE=1 timeout 1500 $T --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 > $O/r2y_bench_n8.json 2> $O/r2y_bench_n8.err; echo "n8 rc $?"
timeout 900 $T --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 2 --ct-vmult replicated --no-e2e --no-check > $O/r2y_bench_n8_repl.json 2> $O/r2y_bench_n8_repl.err; echo "n8 replicated rc $?"
timeout 600 $T --master-port 29523 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > $O/r2y_ref_n8.json 2> $O/r2y_ref_n8.err; echo "ref rc $?"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $T4 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 2 > $O/r2y_bench_n4.json 2> $O/r2y_bench_n4.err; echo "n4 rc $?"
grep -a "e2e step" $O/r2y_bench_n8.err | tail -3
tail -c 400 $O/r2y_bench_n8.err
