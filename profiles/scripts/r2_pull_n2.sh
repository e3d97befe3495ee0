#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29531 tests/multi_gpu_check.py > $O/r2ac_mgc.log 2>&1; echo "mgc rc $?"; grep -a "world\|Error\|error" $O/r2ac_mgc.log | tail -8 | cut -c1-300
timeout 900 $T --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 2 --ct-vmult owned-pull --no-e2e > $O/r2ac_bench_n2_pull.json 2> $O/r2ac_bench_n2_pull.err; echo "pull rc $?"
timeout 900 $T --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-check > $O/r2ac_bench_n2_push.json 2> $O/r2ac_bench_n2_push.err; echo "push rc $?"
python - <<'PY'
import json
for f in ("pull","push"):
    d=json.loads([l for l in open("gpurun_out/r2ac_bench_n2_%s.json"%f) if l.startswith("{")][-1])
    print(f, d["ms_per_step"], d["phases_ms_rank0"]["spmv"], (d.get("ct_vmult_check") or {}).get("per_variant"), d["ct_vmult_reduction"][:80])
PY
tail -3 $O/r2ac_bench_n2_pull.err | cut -c1-300
