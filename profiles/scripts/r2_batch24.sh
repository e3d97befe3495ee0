#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_graded.py -q -m gpu 2>&1 | tail -5
timeout 600 python profiles/graded_timing.py 6 1 > $O/r2af_graded.json 2> $O/r2af_graded.err; echo "timing rc $?"; cat $O/r2af_graded.json | cut -c1-1200; tail -3 $O/r2af_graded.err
