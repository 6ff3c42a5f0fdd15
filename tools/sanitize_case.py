#!/usr/bin/env python3
"""Small end-to-end evaluation for compute-sanitizer runs (memcheck /
racecheck): every kernel variant, ragged tile sizes, halo shards, a batch."""
import os
import sys



ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from colloc_fem_code_b200 import backend, families, sharding, synthetic  # noqa


def main():
    for kind, dims, N in (('ml', (2, 1, 2), 300), ('innovation', (4, 2, 7), 131),
                          ('trapezoid', (2, 1, 2), 257)):
        nx, nu, ny = dims
        exp = synthetic.experiment(1, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        h = p.backend.handle
        for mask in (1, 3, 4, 8, 16, 5, 15, 31):
            h.set_dvec(dvec)
            h.set_multipliers(sigma, lam)
            h.eval(mask)
            h.synchronize()
        for rank in range(2):
            ev = sharding.ShardedEvaluator(p, rank, 2)
            ev.set_point(dvec, sigma, lam)
            ev.handle.eval(backend.ALL)
            ev.handle.fetch(backend.JAC)
            ev.handle.close()
        print(kind, dims, N, 'ok', flush=True)


if __name__ == '__main__':
    main()
