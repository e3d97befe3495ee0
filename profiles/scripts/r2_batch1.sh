#!/bin/bash
# round 2, batch 1: tests, N=1 bench, chain-index comparison, one-rank-of-8 sizing, launch list, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2a_gpu.txt 2>&1
free -g >> $O/r2a_gpu.txt; nproc >> $O/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2a_pytest.log
tail -5 $O/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2a_bench_n1.json 2> $O/r2a_bench_n1.err; echo "bench rc $?"
NMGPU_INDEX=chains timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2a_bench_chains.json 2> $O/r2a_bench_chains.err; echo "chains rc $?"
timeout 900 python bench.py --shard-of 8 --steps 2 --warmup 1 --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2a_shard8.json 2> $O/r2a_shard8.err; echo "shard8 rc $?"
timeout 600 python bench.py --e2e-global-vectors --steps 2 --warmup 1 --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2a_e2e_global.json 2> $O/r2a_e2e_global.err; echo "e2e global rc $?"
K='clip_moments3|clip_local|merge_rows|cand_probe|cand_place|brick_|index_insert|entries_count|hash_clear|t_scatter|t_sort|t_count|split_kernel|bg_boxes|spmv_|fill_line_of|check_index|DeviceScan|bits_popc|row_ids'
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin"
timeout 600 python bench.py $A > $O/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 120 --csv --log-file $O/r2a_launches.csv python bench.py $A > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin"
K6='clip_moments3|merge_rows_fast|cand_probe|brick_insert|t_scatter|t_sort_rows'
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K6" -s 6 -c 6 -o $O/r2a_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
A2="--only-sub-config flower --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-dropin"
K2='clip_local|merge_rows_fast|cand_probe|brick_insert|t_scatter|t_sort_rows|spmv_'
timeout 600 python bench.py $A2 > $O/plain2d.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K2" -s 8 -c 8 -o $O/r2a_prof2d -f python bench.py $A2 > $O/ncu_prof2d.log 2>&1; echo "prof2d rc $?"
ls -la $O | tail -20
