#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x > $O/r2ae_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2ae_pytest.log
NMGPU_BENCH_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2ae_bench.json 2> $O/r2ae_bench.err; echo "bench rc $?"
grep -a "e2e step" $O/r2ae_bench.err | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2ae_bench.json") if l.startswith("{")][-1])
e=d["e2e"]; print(d["ms_per_step"], e["ms_per_step"], e["value"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], e["api"][:160])
PY
