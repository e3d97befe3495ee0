#!/bin/bash
# cand_place_brick bounded to 32 registers (8 blocks per SM instead of 4): candidate tests, then the phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -3
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2av_bench.json 2> $O/r2av_bench.err; echo "bench rc $?"
python - "$O/r2av_bench.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f"%d["ms_per_step"], {k: round(v,3) for k,v in p.items()})
for k,c in d.get("configs",{}).items(): print("   ",k, round(c["ms_per_step"],3), round(c["phases_ms"]["candidates"],3))
PY
