#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_lm_poisson.py tests/test_gpu_schur.py -m gpu -q > $O/r2s_pytest.log 2>&1; echo "pytest rc $?"; tail -30 $O/r2s_pytest.log
