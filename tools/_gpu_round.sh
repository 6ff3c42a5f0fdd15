set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1j.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/pytest_gpu_r1j.log
for mode in peer peer_sync nccl; do
CFEM_REDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_$mode.json 2> gpurun_out/b2_$mode.err; echo "bench $mode rc=$?"
cat gpurun_out/b2_$mode.json
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/b1_r1j.json 2> gpurun_out/b1_r1j.err; echo "bench1 rc=$?"; cat gpurun_out/b1_r1j.json
