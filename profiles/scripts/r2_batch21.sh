#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_dropin_cpp.py tests/test_gpu_codim0.py -m gpu -q > $O/r2x_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/r2x_pytest.log | cut -c1-200
timeout 600 python profiles/codim0_timing.py 8 1600 > $O/r2x_codim0.json 2> $O/r2x_codim0.err; echo "timing rc $?"; cat $O/r2x_codim0.json | cut -c1-900; tail -3 $O/r2x_codim0.err
