#!/bin/bash
# final 1-GPU verification of the round-2 code: tests, smoke, bench line, 1-GPU rows of the scaling table, ncu
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export CFEM_KEEP_STALE=1
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/f_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
timeout 200 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"
run() { local name=$1; shift
  timeout 100 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/f_${name}.json 2> gpurun_out/f_${name}.err
  echo "$name rc=$?"; }
run balanced533 --kind balanced --dims 533
run ndisc427 --kind ndisc_zoh --dims 427
run strong1e7 --n-total 10000000
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/f_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 150 ncu --set full --import-source on --clock-control none -k regex:cfem_sample_kernel_m31 -s 6 -c 2 \
  -f -o gpurun_out/f_sample_kernel_m31 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/f_ncu_full.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json
for n in ('bench', 'balanced533', 'ndisc427', 'strong1e7'):
    try:
        d = json.loads(open(f'gpurun_out/f_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value %.0f ms %.4f K1 %.4f frac %.3f e2e %s api %s pinned %s ok %s' % (
            d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'],
            d['e2e'].get('value'), d.get('e2e_solver_api', {}).get('value'),
            d.get('e2e_solver_api_pinned', {}).get('value'), d['reduce_check']['ok']))
    except Exception as e:
        print(n, 'unreadable', e)
PY
