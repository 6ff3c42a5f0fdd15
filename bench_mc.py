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


def cpu_fit(args):
    """One problem through the same attempts with CPU-oracle callbacks (the
    ``cpu_baseline`` of the Monte-Carlo metric; test infrastructure only)."""
    seed, samples, tol, max_iter = args
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from colloc_fem_code_b200 import fit, nlp, synthetic
    from oracle import ref_models
    from nlp_helpers import OracleEvaluator
    t0 = time.perf_counter()
    p, dec0 = make_case(seed, samples)
    exp = synthetic.experiment(seed, 2 * samples, 5, 3, 3)
    o = ref_models.make_problem('ml_balanced', exp['y'][samples:],
                                exp['u'][samples:], 5)
    db, cb, scaling = fit.ml_setup(p)
    info = None
    for k, att in enumerate(fit.MC_ATTEMPTS):
        if att.get('staged'):
            continue            # the single-stage attempts only
        bounds = fit.free_factor_signs(p, db) if att.get('free_factor_signs') \
            else db
        s = nlp.InteriorPointSolver(OracleEvaluator(o), bounds, cb)
        s.add_num_option('tol', tol)
        s.add_int_option('max_iter', max_iter)
        for key, value in att['options'].items():
            s.add_num_option(key, value)
        s.set_scaling(*scaling)
        _, info = s.solve(dec0)
        info['attempt'] = k
        if fit.solved(info):
            break
    return {'seed': seed, 'status': info['status'], 'attempt': info['attempt'],
            'iterations': info['iterations'],
            'seconds': time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--problems', type=int, default=32,
                    help='problems per GPU')
    ap.add_argument('--samples', type=int, default=250)
    ap.add_argument('--tol', type=float, default=1e-6)
    ap.add_argument('--max-iter', type=int, default=400)
    ap.add_argument('--workers', type=int, default=-1,
                    help='host worker processes per rank (0: threads in one '
                         'process; -1: host cores / ranks - 1)')
    ap.add_argument('--attempts', type=int, default=len_attempts(),
                    help='attempts of the fit procedure (fit.MC_ATTEMPTS)')
    ap.add_argument('--cpu-subset', type=int, default=0,
                    help='rank 0, afterwards: the first K problems again with '
                         'CPU-oracle callbacks on all host cores')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from colloc_fem_code_b200 import fit
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench_mc.py needs a CUDA device')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')     # results only; no data-path collective
    cores = os.cpu_count() or 1
    workers = args.workers if args.workers >= 0 else max(1, cores // world - 1)
    t_setup = time.perf_counter()
    seeds = [rank * args.problems + i for i in range(args.problems)]
    cases = [make_case(sd, args.samples) for sd in seeds]
    problems = [c[0] for c in cases]
    db, cb, scaling = fit.ml_setup(problems[0])
    setup_s = time.perf_counter() - t_setup

    def make_fitter(sub):
        if workers:
            return fit.ParallelBatchFitter(sub, device=local,
                                           workers=min(workers, len(sub)))
        return fit.BatchFitter(sub, device=local)

    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out, report = fit.fit_with_retries(
        make_fitter, problems, [c[1] for c in cases], db, cb, scaling,
        tol=args.tol, max_iter=args.max_iter,
        attempts=fit.MC_ATTEMPTS[:args.attempts],
        log=lambda rec: print(f'rank {rank}: {rec}', file=sys.stderr,
                              flush=True))
    wall = time.perf_counter() - t0
    infos = [info for _, info in out]
    mine = {
        'rank': rank, 'problems': args.problems, 'wall_s': wall,
        'solved': sum(1 for i in infos if fit.solved(i)),
        'statuses': sorted({i['status'] for i in infos}),
        'unsolved_seeds': [sd for sd, i in zip(seeds, infos)
                           if not fit.solved(i)],
        'solved_by_attempt': [sum(1 for i in infos if fit.solved(i)
                                  and i['attempt'] == k)
                              for k in range(args.attempts)],
        'attempts': report, 'setup_s': setup_s,
        'seconds_gpu_callbacks': sum(r['seconds_gpu_callbacks'] or 0.0
                                     for r in report),
        'seconds_kkt_host_sum': sum(i['seconds_kkt'] for i in infos),
        'batched_launch_rounds': sum(r['launch_rounds'] or 0 for r in report),
        'callback_requests': sum(i['callback_calls'] for i in infos),
        'iterations_mean': float(np.mean([i['iterations'] for i in infos])),
    }
    ranks = [mine]
    if world > 1:
        ranks = [None] * world
        dist.all_gather_object(ranks, mine)
    if rank == 0:
        total = sum(r['problems'] for r in ranks)
        nsolved = sum(r['solved'] for r in ranks)
        wall_max = max(r['wall_s'] for r in ranks)
        rec = {
            'metric': 'mc_fits_per_hour', 'unit': 'solved fits/hour',
            'value': nsolved / wall_max * 3600.0, 'n_gpus': world,
            'problems_total': total, 'problems_per_gpu': args.problems,
            'solved': nsolved, 'success_rate': nsolved / total,
            'wall_s': wall_max,
            'solved_by_attempt': [sum(r['solved_by_attempt'][k]
                                      for r in ranks)
                                  for k in range(args.attempts)],
            'attempt_names': [a['name'] for a in
                              fit.MC_ATTEMPTS[:args.attempts]],
            'seconds_gpu_callbacks_max_rank': max(r['seconds_gpu_callbacks']
                                                  for r in ranks),
            'seconds_kkt_host_sum': sum(r['seconds_kkt_host_sum']
                                        for r in ranks),
            'batched_launch_rounds_max_rank': max(r['batched_launch_rounds']
                                                  for r in ranks),
            'host_cores': cores, 'host_worker_processes_per_rank': workers,
            'per_rank': ranks,
            'config': {'workload': 'mc_blackbox_cfem: ML+Balanced (5,3,3), '
                                   f'N={args.samples}, seeds 0..{total - 1}, '
                                   f'{args.problems} problems per GPU, '
                                   'lock-step batched callbacks, whole '
                                   'problems per GPU, no collective',
                       'solver': 'builtin-ipm (no IPOPT in the image); KKT '
                                 'factorisations on the host',
                       'tol': args.tol, 'max_iter': args.max_iter},
        }
        if args.cpu_subset:
            import multiprocessing as mp
            k = min(args.cpu_subset, args.problems)
            t1 = time.perf_counter()
            with mp.get_context('spawn').Pool(min(cores, k)) as pool:
                cpu = pool.map(cpu_fit, [(sd, args.samples, args.tol,
                                          args.max_iter) for sd in seeds[:k]])
            cpu_wall = time.perf_counter() - t1
            cpu_solved = sum(1 for c in cpu if c['status'].startswith('solved'))
            rec['cpu_baseline'] = {
                'value': cpu_solved / cpu_wall * 3600.0,
                'unit': 'solved fits/hour', 'cores': min(cores, k),
                'kind': 'port',
                'sample': f'seeds 0..{k - 1}: the same problems, starts and '
                          'single-stage attempts with CPU-oracle callbacks, '
                          'one process per problem',
                'solved': cpu_solved, 'problems': k, 'wall_s': cpu_wall,
                'seconds_per_problem_mean': float(np.mean(
                    [c['seconds'] for c in cpu])),
                'same_outcome_as_gpu_path': sum(
                    1 for c, i in zip(cpu, infos)
                    if c['status'].startswith('solved') == fit.solved(i))}
        print(json.dumps(rec))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def len_attempts():
    from colloc_fem_code_b200 import fit
    return len(fit.MC_ATTEMPTS)


if __name__ == '__main__':
    main()
