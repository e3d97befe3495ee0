#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_codim0.py -m gpu -q > $O/r2w_pytest.log 2>&1; echo "pytest rc $?"; tail -40 $O/r2w_pytest.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fe.py -m gpu -q -x > $O/r2w_pytest2.log 2>&1; echo "pytest2 rc $?"; tail -3 $O/r2w_pytest2.log
