#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2i_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2i_pytest.log
tail -8 $O/r2i_pytest.log
python __graft_entry__.py smoke > $O/r2i_smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/r2i_smoke.log
