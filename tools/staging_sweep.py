#!/usr/bin/env python3
"""Throughput of the threaded staging path for PAGEABLE solver arrays
(csrc/cfem_host.inl: staged_d2h / staged_h2d) over worker threads x chunk size,
next to the same transfers with page-locked arrays.

    python tools/staging_sweep.py > gpurun_out/staging_sweep.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402


def main():
    nx, nu, ny, N = 2, 1, 2, 1_000_000
    exp = synthetic.experiment(0, N, nx, nu, ny)
    p = families.make_problem('ml', exp['y'], exp['u'], nx)
    st = p.structure
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    lib = backend.Library.for_structure(st)
    jac = np.empty(p.nnzjac)
    x = np.array(dvec)
    combos = [(t, c) for t in (2, 4, 8, 12, 16) for c in (4, 8, 16, 32)]
    for threads, chunk in combos + [('pinned', 0)]:
        if threads != 'pinned':
            os.environ['CFEM_COPY_THREADS'] = str(threads)
            os.environ['CFEM_COPY_CHUNK_MB'] = str(chunk)
        h = backend.Handle(lib, st.N, [d['source'] for d in st.data],
                           st.scalar_values)
        if threads == 'pinned':
            lib.cfem_host_register(jac.ctypes.data, jac.nbytes)
            lib.cfem_host_register(x.ctypes.data, x.nbytes)
        h.set_dvec(x)
        h.set_multipliers(sigma, lam)
        h.eval(backend.ALL)
        h.fetch(backend.JAC, jac)
        d2h, h2d = [], []
        for _ in range(8):
            t0 = time.perf_counter()
            h.fetch(backend.JAC, jac)
            d2h.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            h.set_dvec(x)
            h.synchronize()
            h2d.append(time.perf_counter() - t0)
            h.eval(backend.ALL)         # a new x invalidates the results
        print(json.dumps({
            'threads': threads, 'chunk_mb': chunk,
            'd2h_gbs': jac.nbytes / min(d2h) / 1e9,
            'd2h_gbs_median': jac.nbytes / float(np.median(d2h)) / 1e9,
            'h2d_gbs': x.nbytes / min(h2d) / 1e9,
            'd2h_bytes': jac.nbytes, 'h2d_bytes': x.nbytes,
            'host_cores': os.cpu_count()}), flush=True)
        if threads == 'pinned':
            lib.cfem_host_unregister(jac.ctypes.data)
            lib.cfem_host_unregister(x.ctypes.data)
        h.close()


if __name__ == '__main__':
    main()
