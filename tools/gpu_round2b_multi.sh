#!/bin/bash
# other script shapes time-sharded over NG GPUs + strong scaling
set -u
NG=${1:-2}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CFEM_KEEP_STALE=1
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/m${NG}_gpu.txt 2>&1
run() {   # name, bench args...
  local name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG \
    --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $NG \
    --no-cpu-baseline "$@" > gpurun_out/m${NG}_${name}.json 2> gpurun_out/m${NG}_${name}.err
  echo "$name rc=$? $(tail -c 300 gpurun_out/m${NG}_${name}.json | head -c 10)"
}
run balanced533 --kind balanced --dims 533 --no-e2e
run ndisc427 --kind ndisc_zoh --dims 427 --no-e2e
run strong1e7 --n-total 10000000 --no-e2e
if [ "$NG" = "2" ]; then
  run ml212 
  timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/m${NG}_pytest_multi.log 2>&1
  echo "pytest multi rc=$?"; tail -2 gpurun_out/m${NG}_pytest_multi.log
fi
for f in gpurun_out/m${NG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], d['scaling'], d['config']['family'], d['config']['dims'], d['config']['n_samples_total'],
          'value %.0f sync %s ms %.4f frac %.3f ok %s' % (d['value'], d.get('value_sync'), d['ms_per_step'], d['roofline']['frac'], d['reduce_check']['ok']))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
