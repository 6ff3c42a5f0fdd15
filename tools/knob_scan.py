#!/usr/bin/env python3
"""Scan a run-time knob (environment variable read at cfem_create) and report
the fused kernel's time:  python tools/knob_scan.py CFEM_PREFETCH 0 -1 1000"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402


def main():
    knob, values = sys.argv[1], sys.argv[2:]
    for kind, dims, N in (('ml', (2, 1, 2), 1_000_000),
                          ('balanced', (5, 3, 3), 1_000_000)):
        nx, nu, ny = dims
        exp = synthetic.experiment(0, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx)
        st = p.structure
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        lib = backend.Library.for_structure(st)
        for v in values:
            os.environ[knob] = v
            h = backend.Handle(lib, st.N, [d['source'] for d in st.data],
                               st.scalar_values)
            h.set_kernel_timing(True)
            h.set_dvec(dvec)
            h.set_multipliers(sigma, lam)
            dptr = h.device_ptrs()['dvec']
            for i in range(25):
                h.flush_l2(256 << 20)
                h.set_dvec_device(dptr)
                h.eval(31)
            ms = h.sample_kernel_ms_history(20)
            print(kind, dims, knob, v, 'min %.4f med %.4f' %
                  (min(ms), float(np.median(ms))), flush=True)
            h.close()


if __name__ == '__main__':
    main()
