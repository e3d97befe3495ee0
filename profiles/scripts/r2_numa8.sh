#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi topo -m > $O/r2v_topo.txt 2>&1
lscpu | head -30 >> $O/r2v_topo.txt 2>&1
cat /sys/fs/cgroup/cpuset.mems.effective /sys/fs/cgroup/cpuset.cpus.effective >> $O/r2v_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 profiles/numa_probe.py > $O/r2v_numa.log 2> $O/r2v_numa.err; echo "probe rc $?"
cat $O/r2v_numa.log | cut -c1-400
