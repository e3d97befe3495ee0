#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python profiles/debug_fe33.py > $O/r2e_debug.log 2>&1; cat $O/r2e_debug.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2e_pytest.log
tail -6 $O/r2e_pytest.log
NMGPU_BENCH_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2e_bench_n1.json 2> $O/r2e_bench_n1.err; echo "bench rc $?"
grep "e2e step" $O/r2e_bench_n1.err | head -4
