#!/bin/bash
# ncu --set full of the merge and the two transpose kernels (sphere cycle 6) after the round's late changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
K6='merge_rows_fast|t_scatter32|t_sort_rows32'
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K6" -c 3 -o $O/r2aj_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
ls -la $O/r2aj_prof6.ncu-rep
