set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python tools/roofline_table.py > gpurun_out/roofline_table.jsonl 2> gpurun_out/roofline_table.err; echo "table rc=$?"; cat gpurun_out/roofline_table.jsonl
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1q.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cfem_sample_kernel_m31 -s 4 -c 2 -f -o gpurun_out/prof_r1q python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
cat gpurun_out/plain.log | tail -n 2
