#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
echo skip tests
timeout 600 python profiles/graded_timing.py 6 1 > $O/r2af_graded.json 2> $O/r2af_graded.err; echo "timing rc $?"; cat $O/r2af_graded.json | cut -c1-1200; tail -3 $O/r2af_graded.err
