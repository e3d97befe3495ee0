#!/bin/bash
# round 2, batch 2: brick index v2 (single origin, int math), deterministic merge, merge grouping; PCIe duplex check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2b_pytest.log
tail -5 $O/r2b_pytest.log
python profiles/pcie_duplex.py > $O/r2b_pcie.json 2>&1; cat $O/r2b_pcie.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; echo "bench rc $?"
python profiles/determinism_check.py > $O/r2b_determinism.log 2>&1; echo "determinism rc $?"; tail -3 $O/r2b_determinism.log
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin"
K6='merge_rows_fast|cand_probe|brick_insert|brick_base|brick_fill'
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K6" -s 5 -c 5 -o $O/r2b_prof6 -f python bench.py $A6 > $O/ncu_prof6.log 2>&1; echo "prof6 rc $?"
