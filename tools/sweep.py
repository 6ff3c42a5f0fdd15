#!/usr/bin/env python3
"""Kernel tuning sweep: build variants of one model library (tile size, output
pass budget, ...) and time the fused per-sample kernel on the GPU.

    python tools/sweep.py build   # here (no GPU): compile every variant
    python tools/sweep.py run     # on the GPU box: time them
"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402

KIND = os.environ.get('SWEEP_KIND', 'ml')
DIMS = tuple(int(c) for c in os.environ.get('SWEEP_DIMS', '212'))
N = int(os.environ.get('SWEEP_N', 1_000_000))
VARIANTS = []
if os.environ.get('SWEEP_SET', 'occupancy') == 'store':
    for st_ in (None, 'wb', 'cg'):
        VARIANTS.append(dict(tile=128, pass_budget=6, store=st_))
        VARIANTS.append(dict(tile=128, pass_budget=6, store=st_,
                             experiment='noload'))
elif os.environ.get('SWEEP_SET') == 'r2':
    # round 2: balanced persistent schedule (waves = 1, 2) against one tile
    # per CTA (waves = 8), tile size x launch bounds
    for tile, mbs in ((128, (None, 8)), (64, (None, 16)), (256, (None, 4))):
        for mb in mbs:
            VARIANTS.append(dict(tile=tile, pass_budget=6, min_blocks=mb))
elif os.environ.get('SWEEP_SET') == 'retire':
    # round 2, second session: retirement of the partial sums ("last block
    # done" tree against self-validating slots + finaliser CTA) x first
    # tile's loads before / after the prologue, then tile size, output-pass
    # budget and forced residency on top of the fence-free form
    for retire in ('tree', 'flag'):
        for early in (0, 1):
            VARIANTS.append(dict(tile=128, pass_budget=6, retire=retire,
                                 early_loads=early))
    if os.environ.get('SWEEP_WIDE', '1') == '1':
        for tile in (64, 256):
            VARIANTS.append(dict(tile=tile, pass_budget=6, retire='flag',
                                 early_loads=1))
        for pb in (4, 8, 12):
            VARIANTS.append(dict(tile=128, pass_budget=pb, retire='flag',
                                 early_loads=1))
        for mb in (10, 12):
            VARIANTS.append(dict(tile=128, pass_budget=6, retire='flag',
                                 early_loads=1, min_blocks=mb))
else:       # resident CTAs per SM: launch bounds x single staging buffer
    for tile, mbs in ((64, (None, 16, 18, 20)), (128, (None, 8, 9, 10)),
                      (256, (None, 4, 5))):
        for mb in mbs:
            VARIANTS.append(dict(tile=tile, pass_budget=6, min_blocks=mb))
SINGLE = [int(v) for v in os.environ.get('SWEEP_SINGLE_BUF', '1,0').split(',')]
WAVES = [int(w) for w in os.environ.get('SWEEP_WAVES', '4,8').split(',')]
TAILS = [0]


def problem():
    nx, nu, ny = DIMS
    exp = synthetic.experiment(0, N, nx, nu, ny)
    p = families.make_problem(KIND, exp['y'], exp['u'], nx, dt=0.05)
    return exp, p


def build():
    import concurrent.futures as cf
    nx, nu, ny = DIMS
    p = families.make_problem(KIND, np.zeros((4, ny)), np.zeros((4, nu)), nx,
                              dt=0.05)
    st = p.structure
    with cf.ThreadPoolExecutor(8) as pool:
        futs = [pool.submit(backend.build_library, st,
                            backend.structure_label(st), masks=(31,), **v)
                for v in VARIANTS]
        for v, f in zip(VARIANTS, futs):
            print(v, os.path.basename(f.result()))


def run():
    nx, nu, ny = DIMS
    exp, p = problem()
    st = p.structure
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    balg = 0
    sys.path.insert(0, ROOT)
    import bench
    balg = bench.algorithmic_bytes_per_sample(nx, nu, ny)
    out = []
    for v, waves, single, tail in itertools.product(VARIANTS, WAVES, SINGLE,
                                                    TAILS):
        os.environ['CFEM_WAVES'] = str(waves)
        os.environ['CFEM_TAIL_LEVELS'] = str(tail)
        os.environ['CFEM_SINGLE_BUF'] = str(single)
        lib = backend.Library.load(backend.build_library(
            st, backend.structure_label(st), masks=(31,), **v))
        h = backend.Handle(lib, st.N, [d['source'] for d in st.data],
                           st.scalar_values)
        h.set_kernel_timing(True)
        h.set_dvec(dvec)
        lam_h = lam if h.ncons == lam.size else np.resize(lam, h.ncons)
        h.set_multipliers(sigma, lam_h)
        ms = []
        for i in range(13):
            h.flush_l2(256 << 20)
            h.set_dvec_device(h.device_ptrs()['dvec'])
            h.eval(31)
            ms.append(h.last_sample_kernel_ms())
        ms = ms[3:]
        f_val = float(h.fetch(1)[0])        # sanity: the same objective everywhere
        rec = dict(v, waves=waves, single_buf=single, tail_levels=tail, ms_min=min(ms), ms_med=float(np.median(ms)), f=f_val,
                   gbs=balg * N / (np.median(ms) * 1e-3) / 1e9)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        h.close()
    return out


if __name__ == '__main__':
    {'build': build, 'run': run}[sys.argv[1]]()
