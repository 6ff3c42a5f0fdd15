#!/usr/bin/env python3
"""Run the Monte-Carlo fit procedure (``fit.MC_ATTEMPTS``) on chosen seeds with
CPU-oracle callbacks, one process per seed -- the staged fifth attempt
included.  Used to examine the seeds that a GPU batch left unsolved without
spending GPU time: the interior-point driver produces the same iterates with
oracle and with CUDA callbacks (tests/test_gpu_nlp.py), so which attempt
solves a seed is the same on both.  Test / analysis infrastructure only.

    python tests/mc_cpu_attempts.py --seeds 28,31,34 --from-attempt 4 \
        > profiles/r02_mc_unsolved_seeds_cpu.jsonl
"""
import argparse
import concurrent.futures as cf
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))     # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


class OracleFitter:
    """``fit.BatchFitter`` surface for a list of problems, oracle callbacks,
    problems solved one after the other."""

    def __init__(self, problems):
        from oracle import ref_models
        self.problems = problems
        self.oracles = []
        for p in problems:
            kind = getattr(p, 'family', None) or \
                ('ml_balanced' if 'sPp_tril' in p.decision else 'balanced')
            self.oracles.append(ref_models.make_problem(kind, p.y, p.u,
                                                        p.model.nx))
        self.launches = 0
        self.seconds_gpu = 0.0

    def fit(self, dec0s, dec_bounds, constr_bounds, scaling, tol=1e-8,
            max_iter=400, options=None):
        from colloc_fem_code_b200 import nlp
        from nlp_helpers import OracleEvaluator
        out = []
        for o, d in zip(self.oracles, dec0s):
            s = nlp.InteriorPointSolver(OracleEvaluator(o), dec_bounds,
                                        constr_bounds)
            s.add_num_option('tol', tol)
            s.add_int_option('max_iter', max_iter)
            for key, value in (options or {}).items():
                s.add_num_option(key, value)
            s.set_scaling(*scaling)
            out.append(s.solve(np.asarray(d)))
        return out

    def close(self):
        pass


def run_seed(args):
    seed, samples, tol, max_iter, first = args
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    import bench_mc
    from colloc_fem_code_b200 import fit
    t0 = time.perf_counter()
    p, dec0 = bench_mc.make_case(seed, samples)
    db, cb, scaling = fit.ml_setup(p)
    res, report = fit.fit_with_retries(
        OracleFitter, [p], [dec0], db, cb, scaling, tol=tol,
        max_iter=max_iter, attempts=fit.MC_ATTEMPTS[first:])
    info = res[0][1]
    return {'seed': seed, 'status': info['status'],
            'attempt': fit.MC_ATTEMPTS[first + info['attempt']]['name'],
            'iterations': info['iterations'],
            'stage1_status': info.get('stage1_status'),
            'attempts_run': [r['attempt'] for r in report],
            'seconds': time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', required=True)
    ap.add_argument('--samples', type=int, default=250)
    ap.add_argument('--tol', type=float, default=1e-6)
    ap.add_argument('--max-iter', type=int, default=400)
    ap.add_argument('--from-attempt', type=int, default=0,
                    help='index into fit.MC_ATTEMPTS of the first attempt')
    ap.add_argument('--procs', type=int, default=os.cpu_count())
    a = ap.parse_args()
    seeds = [int(s) for s in a.seeds.split(',')]
    jobs = [(s, a.samples, a.tol, a.max_iter, a.from_attempt) for s in seeds]
    with cf.ProcessPoolExecutor(min(a.procs, len(jobs))) as pool:
        for rec in pool.map(run_seed, jobs):
            print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
