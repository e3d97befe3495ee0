#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x > $O/r2r_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2r_pytest.log
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2r_bench.json 2> $O/r2r_bench.err; echo "rc $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2r_bench.json") if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f clip %.3f merge %.2f k_merge %.3f"%(d["ms_per_step"],p["k_clip"],p["merge"],p["k_merge"]))
for k,v in d["configs"].items(): print(k, v["value"], v["ms_per_step"])
PY
