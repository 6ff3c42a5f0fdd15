set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nlp.py -m gpu -x -q 2>&1 | tail -n 6
timeout 900 python tools/roofline_table.py > gpurun_out/roofline_table.jsonl 2> gpurun_out/roofline_table.err; echo "table rc=$?"; python -c "
import json
for l in open('gpurun_out/roofline_table.jsonl'):
    d=json.loads(l); print(d['config'], d['dims'], d['N'], round(d['kernel_ms'],4), round(d['step_ms'],4), round(d['frac_of_measured_peak'],3), d.get('e2e_host_api_ms'), d.get('cpu_oracle_ms'))"
