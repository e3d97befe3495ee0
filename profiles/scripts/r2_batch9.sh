#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2k_bench.json 2> $O/r2k_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2k_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["phases_ms_rank0"])
PY
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'clip_moments3' -s 1 -c 1 -o $O/r2k_clip6 -f python bench.py $A6 > $O/ncu_clip6.log 2>&1; echo "prof rc $?"
