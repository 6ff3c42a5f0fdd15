set -x
for n in 8 4; do
CFEM_REDUCE=peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_${n}_final.json 2> gpurun_out/scale_${n}_final.err; echo "bench $n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/scale_${n}_final.json')); print(d['value'], d['ms_per_step'], d['per_rank'], d['e2e']['value'])"
done
