set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_r1l.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_multi_r1l.log
for mode in peer peer_sync; do
CFEM_REDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/b2_$mode.json 2> gpurun_out/b2_$mode.err; echo "bench $mode rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/b2_$mode.json')); print(d['value'], d['ms_per_step'], d['per_rank'], d['e2e']['value'], d['wall_s_timed_region'], d['host_enqueue_s'])"
done
timeout 600 python tools/step_overhead.py > gpurun_out/step_overhead.jsonl 2> gpurun_out/step_overhead.err; cat gpurun_out/step_overhead.jsonl; tail -n 5 gpurun_out/step_overhead.err
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -n 3
