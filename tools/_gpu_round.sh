set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_gpu_final2.log
for mode in peer; do
CFEM_REDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/b2_$mode.json 2> gpurun_out/b2_$mode.err; echo "bench $mode rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/b2_$mode.json')); print(d['value'], d['ms_per_step'], d['per_rank'], d['e2e']['value'], d['gpu_launches'])"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 | cut -c1-400
