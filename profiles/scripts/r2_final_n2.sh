#!/bin/bash
# round 2: 2 GPUs -- multi-GPU test, bench with checks (owned default), replicated variant, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r2y_gpu.txt; free -g >> $O/r2y_gpu.txt; nproc >> $O/r2y_gpu.txt
timeout 900 python -m pytest tests/test_gpu_round2.py -k two_gpus -q > $O/r2y_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2y_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NMGPU_BENCH_TRACE=1 timeout 1200 $T --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2y_bench_n2.json 2> $O/r2y_bench_n2.err; echo "n2 rc $?"
timeout 900 $T --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --ct-vmult replicated --no-e2e > $O/r2y_bench_n2_repl.json 2> $O/r2y_bench_n2_repl.err; echo "n2 replicated rc $?"
timeout 900 $T --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 --nccl --no-e2e --no-check > $O/r2y_bench_n2_nccl.json 2> $O/r2y_bench_n2_nccl.err; echo "n2 nccl rc $?"
timeout 600 $T --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r2y_ref_n2.json 2> $O/r2y_ref_n2.err; echo "ref rc $?"
tail -c 600 $O/r2y_bench_n2.err
