#!/bin/bash
# 32-bit offsets written by the scan itself (no separate pass): whole GPU suite, then the phases of the benchmark step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2as_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2as_pytest.log
tail -8 $O/r2as_pytest.log
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2as_bench.json 2> $O/r2as_bench.err; echo "bench rc $?"
python - "$O/r2as_bench.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f"%d["ms_per_step"], {k: round(v,3) for k,v in p.items()})
PY
