#!/bin/bash
# whole GPU suite after the merge / transpose changes, then the phases of the benchmark step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2am_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2am_pytest.log
tail -8 $O/r2am_pytest.log
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2am_bench.json 2> $O/r2am_bench.err; echo "bench rc $?"
python - "$O/r2am_bench.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f"%d["ms_per_step"], {k: round(v,3) for k,v in p.items()})
PY
