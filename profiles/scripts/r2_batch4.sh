#!/bin/bash
# round 2, batch 4: general-FE path tests, async read-back probe, bench after reverting to the match-based merge
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2d_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2d_pytest.log
tail -15 $O/r2d_pytest.log
timeout 300 python profiles/async_probe.py > $O/r2d_async.log 2>&1; cat $O/r2d_async.log
NMGPU_BENCH_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2d_bench_n1.json 2> $O/r2d_bench_n1.err; echo "bench rc $?"
grep "e2e step" $O/r2d_bench_n1.err | tail -2
