#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_schur.py -m gpu -q > $O/r2j_schur.log 2>&1; echo "pytest rc $?" >> $O/r2j_schur.log
tail -8 $O/r2j_schur.log
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _g1 _g2 _g4; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so NMGPU_BENCH_TRACE=1 timeout 300 python bench.py $B > $O/r2j_bench$v.json 2> $O/r2j_bench$v.err
  echo "variant '$v' rc $?"; tail -c 1500 $O/r2j_bench$v.json
done
