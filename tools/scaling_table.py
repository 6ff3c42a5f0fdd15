#!/usr/bin/env python3
"""Collect bench.py lines of several GPU counts / workloads into one JSON-lines
table (profiles/r02_scaling_table.jsonl) and print it as Markdown.

    python tools/scaling_table.py out.jsonl line1.json line2.json ...
"""
import json
import sys


def row(path):
    with open(path) as fh:
        d = json.loads(fh.read().strip().splitlines()[-1])
    cfg = d['config']
    r = {
        'file': path.split('/')[-1], 'family': cfg['family'],
        'dims': cfg['dims'], 'n_gpus': d['n_gpus'], 'scaling': d['scaling'],
        'n_samples_total': cfg['n_samples_total'],
        'sets_per_s': d['value'], 'ms_per_step': d['ms_per_step'],
        'sets_per_s_sync': d.get('value_sync'),
        'ms_per_step_sync': d.get('ms_per_step_sync'),
        'samples_per_s': d['samples_per_s'],
        'kernel_ms_per_rank': d['per_rank']['kernel_ms'],
        'roofline_frac_rank0': d['roofline']['frac'],
        'reduce_check_ok': d['reduce_check']['ok'],
        'reduce_check_rel_err': d['reduce_check']['rel_err'],
        'same_bits_on_all_ranks': d['reduce_check']['same_bits_on_all_ranks'],
        'e2e_sets_per_s': d['e2e'].get('value'),
        'e2e_ms_median': d['e2e'].get('ms_per_step_median'),
        'unit': d['unit'],
    }
    return r


def main():
    out, files = sys.argv[1], sys.argv[2:]
    rows = [row(f) for f in files]
    rows.sort(key=lambda r: (r['scaling'], r['family'], r['n_gpus']))
    with open(out, 'w') as fh:
        for r in rows:
            fh.write(json.dumps(r) + '\n')
    base = {}
    for r in rows:
        if r['n_gpus'] == 1:
            base[(r['scaling'], r['family'], tuple(r['dims']))] = r
    print('| workload | scaling | GPUs | N total | sets/s | ms/step | '
          'sync sets/s | efficiency | sync eff. | K1 frac | reduce_check |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for r in rows:
        b = base.get((r['scaling'], r['family'], tuple(r['dims'])))
        eff = seff = ''
        if b:
            ideal = b['sets_per_s'] * r['n_gpus']
            eff = f"{r['sets_per_s'] / ideal:.3f}"
            if r['sets_per_s_sync']:
                seff = f"{r['sets_per_s_sync'] / ideal:.3f}"
        sync = f"{r['sets_per_s_sync']:.0f}" if r['sets_per_s_sync'] else ''
        print(f"| {r['family']} {tuple(r['dims'])} | {r['scaling']} | "
              f"{r['n_gpus']} | {r['n_samples_total']:,} | "
              f"{r['sets_per_s']:.0f} | {r['ms_per_step']:.4f} | {sync} | "
              f"{eff} | {seff} | {r['roofline_frac_rank0']:.3f} | "
              f"{'ok' if r['reduce_check_ok'] else 'FAIL'} "
              f"({r['reduce_check_rel_err']:.1e}) |")


if __name__ == '__main__':
    main()
