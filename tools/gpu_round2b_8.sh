#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
  --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 \
  --no-cpu-baseline --n-total 10000000 --no-e2e > gpurun_out/m8_strong1e7.json 2> gpurun_out/m8_strong1e7.err
echo "strong1e7 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/m8_strong1e7.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], d['config']['n_samples_total'], 'value %.0f sync %s ms %.4f frac %.3f ok %s' % (
    d['value'], d.get('value_sync'), d['ms_per_step'], d['roofline']['frac'], d['reduce_check']['ok']))
PY
