set -x
SWEEP_WAVES=8,16 timeout 900 python tools/sweep.py run > gpurun_out/sweep6_212.log 2> gpurun_out/sweep6_212.err; tail -n 3 gpurun_out/sweep6_212.err; cat gpurun_out/sweep6_212.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -n 3
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/b1_r1n.json 2> gpurun_out/b1_r1n.err; echo "bench1 rc=$?"; cat gpurun_out/b1_r1n.json
