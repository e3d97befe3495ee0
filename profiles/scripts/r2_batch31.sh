#!/bin/bash
# SpMV padding slots read the row's first column instead of entry 0: regression tests, then the product timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_fe.py tests/test_gpu_parity.py tests/test_gpu_schur.py -q -m gpu 2>&1 | tail -6
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 300 python bench.py $B > $O/r2an_bench.json 2> $O/r2an_bench.err; echo "bench rc $?"
python - "$O/r2an_bench.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f"%d["ms_per_step"], {k: round(v,3) for k,v in p.items()}); print(d["spmv"]["Ct_vmult_frac_hbm"], d["spmv"]["C_Tvmult_frac_hbm"])
PY
