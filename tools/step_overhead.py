#!/usr/bin/env python3
"""Where does the time of one callback set go beyond the per-sample kernel?

Times back-to-back evaluations (events on the handle's stream around every
step, L2 flushed outside the events) with and without the per-kernel timing
events and with the parameter-only kernel (and its fork/join onto the
auxiliary stream) switched off (CFEM_SKIP_PARAM, measurement only)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402


def measure(lib, st, dvec, lam, sigma, timing, steps=40):
    h = backend.Handle(lib, st.N, [d['source'] for d in st.data],
                       st.scalar_values)
    h.set_kernel_timing(timing)
    h.set_dvec(dvec)
    h.set_multipliers(sigma, lam)
    dptr = h.device_ptrs()['dvec']
    ms = []
    for i in range(steps + 5):
        h.flush_l2(256 << 20)
        h.event_record(0)
        h.set_dvec_device(dptr)
        h.eval(31)
        h.event_record(1)
        h.synchronize()
        if i >= 5:
            ms.append(h.event_elapsed_ms(0, 1))
    k = h.sample_kernel_ms_history(20) if timing else [float('nan')]
    h.close()
    return float(np.median(ms)), float(np.min(ms)), float(np.median(k))


def main():
    for kind, dims, N in (('ml', (2, 1, 2), 1_000_000),
                          ('ml', (2, 1, 2), 1_000),
                          ('balanced', (5, 3, 3), 250)):
        nx, nu, ny = dims
        exp = synthetic.experiment(0, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx)
        st = p.structure
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        lib = backend.Library.for_structure(st)
        for label, env, timing in (('default+timing', {}, True),
                                   ('default', {}, False),
                                   ('pdl', {'CFEM_PDL': '2'}, False),
                                   ('fork-join', {'CFEM_PDL': '0'}, False),
                                   ('no param kernel', {'CFEM_SKIP_PARAM': '1'},
                                    False),
                                   ('graph', {'CFEM_PDL': '0', 'CFEM_GRAPH': '1'}, False)):
            for k, v in env.items():
                os.environ[k] = v
            med, mn, kms = measure(lib, st, dvec, lam, sigma, timing)
            for k in env:
                del os.environ[k]
            print(json.dumps({'kind': kind, 'dims': dims, 'N': N,
                              'variant': label, 'step_ms_median': med,
                              'step_ms_min': mn, 'k1_ms_median': kms}),
                  flush=True)


if __name__ == '__main__':
    main()
