#!/usr/bin/env python3
"""A/B of the two-kernel launch forms at the scripts' native lengths:
alternating rounds, many steps, medians (tools/step_overhead.py is one short
round per variant)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402
from step_overhead import measure  # noqa: E402

VARIANTS = (('default', {}), ('pdl', {'CFEM_PDL': '2'}), ('fork-join', {'CFEM_PDL': '0'}),
            ('graph', {'CFEM_PDL': '0', 'CFEM_GRAPH': '1'}),
            ('no param kernel', {'CFEM_SKIP_PARAM': '1'}))


def main():
    for kind, dims, N in (('ml', (2, 1, 2), 1_000),
                          ('balanced', (5, 3, 3), 250),
                          ('ndisc_zoh', (4, 2, 7), 601),
                          ('ml', (2, 1, 2), 1_000_000)):
        nx, nu, ny = dims
        exp = synthetic.experiment(0, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
        st = p.structure
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        lib = backend.Library.for_structure(st)
        res = {v[0]: [] for v in VARIANTS}
        for rnd in range(4):
            for label, env in VARIANTS:
                for k, v in env.items():
                    os.environ[k] = v
                med, mn, _ = measure(lib, st, dvec, lam, sigma, False,
                                     steps=150 if N < 10_000 else 40)
                for k in env:
                    del os.environ[k]
                res[label].append(med)
        print(json.dumps({'kind': kind, 'dims': dims, 'N': N,
                          'step_us_median_per_round':
                          {k: [round(1e3 * x, 2) for x in v]
                           for k, v in res.items()}}), flush=True)


if __name__ == '__main__':
    main()
