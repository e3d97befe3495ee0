#!/bin/bash
# round 2, final single-GPU pass: tests, smoke, default bench + reference arm, launch list, ncu captures (3D top kernels, 2D)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2z_gpu.txt 2>&1; free -g >> $O/r2z_gpu.txt; nproc >> $O/r2z_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/r2z_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2z_pytest.log
tail -6 $O/r2z_pytest.log
python __graft_entry__.py smoke > $O/r2z_smoke.log 2>&1; echo "smoke rc $?"; tail -2 $O/r2z_smoke.log
timeout 1200 python bench.py > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $O/r2z_ref_n1.json 2> $O/r2z_ref_n1.err; echo "ref rc $?"
K='clip_moments3|clip_local|merge_rows|cand_probe|cand_place|brick_|index_insert|entries_count|hash_clear|t_scatter|t_sort|t_count|split_kernel|bg_boxes|spmv_|fill_line_of|check_|DeviceScan|bits_popc|row_ids|narrow'
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
timeout 600 python bench.py $A > $O/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 160 --csv --log-file $O/r2z_launches.csv python bench.py $A > $O/ncu_launch.log 2>&1; echo "launch list rc $?"
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
K6='clip_moments3|merge_rows_fast|cand_probe|brick_insert|t_scatter|t_sort_rows|spmv_'
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K6" -s 8 -c 8 -o $O/r2z_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
ls -la $O | tail -12
