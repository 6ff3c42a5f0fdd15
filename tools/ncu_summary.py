#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into profiles/<name>.json + traffic.json.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_sample_kernel
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
    'launch__shared_mem_per_block_dynamic',
    'launch__shared_mem_per_block_static',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__cycles_active.avg',
]


def to_bytes(val, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(val) * scale.get(unit, 1)


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        rec = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[k] = {'value': float(r[i].replace(',', '')),
                              'unit': units[i]}
                except ValueError:
                    rec[k] = {'value': r[i], 'unit': units[i]}
        launches.append(rec)
    with open(out + '.json', 'w') as fh:
        json.dump({'report': os.path.basename(rep), 'launches': launches}, fh,
                  indent=1)
    traffic = {}
    for rec in launches:
        name = rec['kernel'].split('(')[0]
        rd, wr = rec['dram__bytes_read.sum'], rec['dram__bytes_write.sum']
        b = to_bytes(rd['value'], rd['unit']) + to_bytes(wr['value'],
                                                         wr['unit'])
        traffic.setdefault(name, []).append(b)
    tpath = os.path.join(os.path.dirname(out), 'traffic.json')
    with open(tpath, 'w') as fh:
        json.dump({k: {'dram_bytes': sum(v) / len(v), 'launches': len(v),
                       'source': os.path.basename(out) + '.json'}
                   for k, v in traffic.items()}, fh, indent=1)
    for rec in launches:
        print(rec['kernel'][:50],
              rec['gpu__time_duration.sum']['value'],
              rec['gpu__time_duration.sum']['unit'])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
