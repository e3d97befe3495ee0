#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2h_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2h_pytest.log
tail -12 $O/r2h_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err; echo "bench rc $?"
