#!/usr/bin/env python3
"""Monte-Carlo batch benchmark (BASELINE.json config 4): independent black-box
identifications, (nx, nu, ny) = (5, 3, 3), ML + balanced-realisation problem
(the composition of /root/reference/mc_blackbox_cfem.py:25-30), N = 250
estimation samples each (the second half of 500, mc_blackbox_cfem.py:40-44),
whole problems per GPU, no collective.

    python bench_mc.py --problems 64 [--gpus N via torchrun]

Every problem runs its own interior-point iteration; the callbacks of all
problems of a round are ONE launch of the fused kernel (blockIdx.y = problem).
The KKT factorisations stay on the host and their time is reported separately.
Prints one JSON line: fits/hour (solved fits only), success rate, time split.
The reference's own MC procedure is a stub (``estimate`` is ``pass``), so there
is no reference number; data and initial guesses are synthetic (true system
perturbed by 10 %, balanced, steady-state Kalman filter start).
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def make_case(seed, N, nx=5, nu=3, ny=3, pert=0.1):
    from colloc_fem_code_b200 import families, fit, synthetic
    exp = synthetic.experiment(seed, 2 * N, nx, nu, ny)
    ue, ye = exp['u'][N:], exp['y'][N:]
    rng = np.random.default_rng(seed + 1)
    A0 = exp['A'] + pert * 0.05 * rng.normal(size=(nx, nx))
    B0 = exp['B'] * (1 + pert * rng.normal(size=(nx, nu)))
    C0 = exp['C'] * (1 + pert * rng.normal(size=(ny, nx)))
    T, bal = fit.balanced_guess(A0, B0, C0)
    g = fit.kalman_guess(ye, ue, bal['A'], bal['B'], bal['C'], exp['D'],
                         0.05 * np.eye(nx), 0.2 * np.eye(ny))
    g.update({k: bal[k] for k in ('sW_diag', 'ctrl_orth', 'obs_orth')})
    p = families.make_problem('ml_balanced', ye, ue, nx)
    return p, fit.start_point(p, g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--problems', type=int, default=32,
                    help='problems per GPU')
    ap.add_argument('--samples', type=int, default=250)
    ap.add_argument('--tol', type=float, default=1e-6)
    ap.add_argument('--max-iter', type=int, default=200)
    ap.add_argument('--workers', type=int, default=0,
                    help='host worker processes (0: threads in one process)')
    args = ap.parse_args()
    import torch
    from colloc_fem_code_b200 import fit
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench_mc.py needs a CUDA device')
    cases = [make_case(rank * args.problems + i, args.samples)
             for i in range(args.problems)]
    problems = [c[0] for c in cases]
    db, cb, scaling = fit.ml_setup(problems[0])
    if args.workers:
        bf = fit.ParallelBatchFitter(problems, device=local,
                                     workers=args.workers)
    else:
        bf = fit.BatchFitter(problems, device=local)
    t0 = time.perf_counter()
    out = bf.fit([c[1] for c in cases], db, cb, scaling, tol=args.tol,
                 max_iter=args.max_iter)
    wall = time.perf_counter() - t0
    solved = sum(1 for _, info in out if info['status'].startswith('solved'))
    rec = {
        'metric': 'mc_fits_per_hour', 'unit': 'solved fits/hour',
        'value': solved / wall * 3600.0, 'rank': rank, 'n_gpus': world,
        'problems': args.problems, 'solved': solved,
        'statuses': sorted({info['status'] for _, info in out}),
        'wall_s': wall, 'seconds_gpu_callbacks': bf.seconds_gpu,
        'seconds_kkt_host_sum': sum(i['seconds_kkt'] for _, i in out),
        'batched_launch_rounds': bf.launches,
        'callback_requests': sum(i['callback_calls'] for _, i in out),
        'iterations_mean': float(np.mean([i['iterations'] for _, i in out])),
        'host_threads': getattr(bf, 'threads', None),
        'host_worker_processes': getattr(bf, 'workers', None),
        'host_cores': os.cpu_count(),
        'config': {'workload': 'mc_blackbox_cfem: ML+Balanced (5,3,3), '
                               f'N={args.samples}, {args.problems} problems '
                               'per GPU, lock-step batched callbacks',
                   'solver': 'builtin-ipm (no IPOPT in the image)',
                   'tol': args.tol, 'max_iter': args.max_iter},
    }
    if hasattr(bf, 'close'):
        bf.close()
    print(json.dumps(rec))


if __name__ == '__main__':
    main()
