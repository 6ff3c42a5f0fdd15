set -x
for n in 8; do
for mode in peer nccl; do
CFEM_REDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_${n}_$mode.json 2> gpurun_out/scale_${n}_$mode.err; echo "bench $n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/scale_${n}_$mode.json')); print(d['value'], d['ms_per_step'], d['per_rank'], d['e2e']['value'], d['wall_s_timed_region'], d['host_enqueue_s'])"
done
done
CUDA_VISIBLE_DEVICES=3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_1_samebox.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/scale_1_samebox.json')); print(d['value'], d['ms_per_step'], d['per_rank'])"
