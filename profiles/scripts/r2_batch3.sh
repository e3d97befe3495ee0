#!/bin/bash
# round 2, batch 3: section-polygon clip kernel vs box-plane clipping, fixed-point merge, cand_probe tweaks, e2e trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c_pytest.log
tail -5 $O/r2c_pytest.log
Q="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin"
timeout 600 python bench.py $Q > $O/r2c_bench_section.json 2> $O/r2c_bench_section.err; echo "section rc $?"
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu_clip.so timeout 600 python bench.py $Q > $O/r2c_bench_clip.json 2> $O/r2c_bench_clip.err; echo "clip rc $?"
NMGPU_BENCH_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2c_bench_e2e.json 2> $O/r2c_bench_e2e.err; echo "e2e rc $?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dropin --no-e2e > $O/r2c_bench_sub.json 2> $O/r2c_bench_sub.err; echo "sub rc $?"
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin"
K6='clip_moments3|merge_rows_fast|cand_probe'
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K6" -s 3 -c 3 -o $O/r2c_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
