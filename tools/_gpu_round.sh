set -x
timeout 600 python bench.py > gpurun_out/b1_final.json 2> gpurun_out/b1_final.err; echo "bench rc=$?"; cat gpurun_out/b1_final.json | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cfem_sample_kernel_m31 -s 4 -c 2 -f -o gpurun_out/prof_final python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
