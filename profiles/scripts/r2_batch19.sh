#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_dropin_cpp.py tests/test_gpu_fe.py tests/test_gpu_lm_poisson.py -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2u_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-check > $O/r2u_bench.json 2> $O/r2u_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2u_bench.json") if l.startswith("{")][-1])
print(d["e2e_dropin"])
PY
