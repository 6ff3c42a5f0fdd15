#!/usr/bin/env python3
"""Warp-stall (PC sampling) breakdown of one kernel of an ``ncu --set full``
report: totals per stall reason and the SASS instructions that collect the
samples, each with the instructions in front of it (a stall is charged to the
instruction that could not issue, so the cause is what precedes it).

    python tools/stall_breakdown.py profiles/r02_sample_kernel_m31.ncu-rep \
        > profiles/r02_stall_breakdown.md

Runs on the CPU (needs only the ncu command-line tool to read the report).
"""
import csv
import io
import subprocess
import sys


def ncu_csv(report, page, extra=()):
    out = subprocess.run(['ncu', '-i', report, '--page', page, '--csv',
                          *extra], capture_output=True, text=True, check=True)
    return list(csv.reader(io.StringIO(out.stdout)))


def main():
    report = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    raw = ncu_csv(report, 'raw')
    hdr, vals = raw[0], raw[2]
    col = {h: v for h, v in zip(hdr, vals)}
    name = col.get('Kernel Name', '?')
    total = int(col['smsp__pcsamp_sample_count'])
    reasons = {}
    for h, v in col.items():
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and \
                not h.endswith('_not_issued'):
            reasons[h[len('smsp__pcsamp_warps_issue_stalled_'):]] = int(v)
    print(f'# Warp-stall breakdown of `{name}`\n')
    print(f'Report `{report}`, first captured launch: '
          f'{float(col["gpu__time_duration.sum"]):.1f} us, '
          f'{total} PC samples '
          f'(interval {col["smsp__pcsamp_interval_cycles"]} cycles), '
          f'{col["launch__registers_per_thread"]} registers, '
          f'grid {col["launch__grid_size"]} x {col["launch__block_size"]}.\n')
    cyc = {k: float(col[f'sm__cycles_active.{k}']) for k in ('avg', 'min', 'max')}
    print(f'SM busy cycles: avg {cyc["avg"]:.0f}, min {cyc["min"]:.0f}, '
          f'max {cyc["max"]:.0f} of '
          f'{float(col["sm__cycles_elapsed.max"]):.0f} elapsed.\n'
          if 'sm__cycles_elapsed.max' in col else
          f'SM busy cycles: avg {cyc["avg"]:.0f}, min {cyc["min"]:.0f}, '
          f'max {cyc["max"]:.0f}.\n')
    print('| stall reason | samples | share |')
    print('|---|---|---|')
    for k, v in sorted(reasons.items(), key=lambda kv: -kv[1]):
        if v:
            print(f'| {k} | {v} | {100.0 * v / total:.1f} % |')
    src = ncu_csv(report, 'source', ('--print-source', 'sass'))
    starts = [i for i, r in enumerate(src) if r and r[0] == 'Kernel Name']
    h = src[starts[0] + 1]
    ix = {n: i for i, n in enumerate(h)}
    end = starts[1] if len(starts) > 1 else len(src)
    body = [r for r in src[starts[0] + 2:end] if len(r) > ix['# Samples']]
    stalls = [n for n in h if n.startswith('stall_') and 'Not' not in n]
    order = sorted(range(len(body)),
                   key=lambda i: -int(body[i][ix['# Samples']] or 0))[:top_n]
    print(f'\n## The {top_n} instructions with the most samples\n')
    print('Each entry: samples, share, stall reasons; then the instructions '
          'in front of it.\n')
    for i in order:
        r = body[i]
        n = int(r[ix['# Samples']] or 0)
        why = ', '.join(f'{s[6:]} {int(r[ix[s]] or 0)}' for s in stalls
                        if int(r[ix[s]] or 0))
        print(f'* **{n}** ({100.0 * n / total:.1f} %) at '
              f'`{r[ix["Source"]].strip()}` — {why}')
        print('  ```')
        for q in body[max(0, i - 4):i + 1]:
            print(f'  {q[ix["Address"]][-5:]}  {q[ix["Source"]].strip()}')
        print('  ```')


if __name__ == '__main__':
    main()
