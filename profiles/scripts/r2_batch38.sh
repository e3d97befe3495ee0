#!/bin/bash
# ncu --set full of cand_place_brick (sphere cycle 6)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 300 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cand_place_brick" -c 1 -o $O/r2au_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
