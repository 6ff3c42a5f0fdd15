#!/bin/bash
# call 1 of the second round-2 session: retirement / prologue variants of K1
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CFEM_KEEP_STALE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
for spec in "ml 212 1" "balanced 533 0" "ndisc_zoh 427 0"; do
  set -- $spec
  SWEEP_SET=retire SWEEP_WIDE=$3 SWEEP_KIND=$1 SWEEP_DIMS=$2 SWEEP_WAVES=8 SWEEP_SINGLE_BUF=1 \
    timeout 120 python tools/sweep.py run > gpurun_out/c1_sweep_$1_$2.jsonl 2> gpurun_out/c1_sweep_$1_$2.err
  echo "sweep $1 $2 rc=$?"
done
# a second pass of the four main variants (run-to-run noise)
SWEEP_SET=retire SWEEP_WIDE=0 SWEEP_KIND=ml SWEEP_DIMS=212 SWEEP_WAVES=8,4 SWEEP_SINGLE_BUF=1 \
  timeout 100 python tools/sweep.py run > gpurun_out/c1_sweep_ml_212_pass2.jsonl 2>> gpurun_out/c1_sweep_ml_212.err
echo "sweep pass2 rc=$?"
CFEM_RETIRE=flag CFEM_EARLY_LOADS=1 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/c1_bench_flag1.json 2> gpurun_out/c1_bench_flag1.err
echo "bench flag1 rc=$?"
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/c1_bench_tree0.json 2> gpurun_out/c1_bench_tree0.err
echo "bench tree0 rc=$?"
CFEM_RETIRE=flag CFEM_EARLY_LOADS=1 timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest_flag1.log 2>&1
echo "pytest flag1 rc=$?"
tail -3 gpurun_out/c1_pytest_flag1.log
cat gpurun_out/c1_sweep_ml_212.jsonl
