#!/usr/bin/env python3
"""Benchmark of the hot path: NLP callback evaluations per second.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one full callback set at a new decision vector -- objective,
gradient, constraints, constraint-Jacobian values and Lagrangian-Hessian values
(IPOPT's eval_f + eval_grad_f + eval_g + eval_jac_g + eval_h), per-sample AND
parameter-only functions -- of the ATTAS short-period maximum-likelihood
problem ((nx, nu, ny) = (2, 1, 2), attas_sp_ml.py:79-87 of the reference) on a
synthetic trajectory of N = 1e6 samples per GPU.  With more than one GPU the
trajectory is N = 1e6 x n_gpus samples, split in time with a one-sample halo;
the objective and the parameter block of the gradient are summed across GPUs
every step INSIDE the per-sample kernel over NVLink peer memory (pipelined:
the kernel of step k+1 finishes the sums of step k; CFEM_REDUCE=peer_sync
makes every kernel wait for its own sums, CFEM_REDUCE=nccl uses an NCCL
all-reduce plus a small kernel instead).  Weak scaling.

``value``  callback sets per second with inputs resident in HBM (events around
           every step on the launching stream, L2 flushed between steps);
           one "set" is normalised to N = 1e6 samples.
``value_sync``  (N > 1) the same with the synchronous exchange: every kernel
           returns with its own reduced objective / gradient, which is what a
           sequential NLP solver needs; ``value`` is the pipelined exchange.
``reduce_check``  after the timed passes every rank compares its objective and
           the sRp-diagonal entries of its gradient with the closed form
           -1/2 sum en^2 - N sum log sRp_ii and -N / sRp_ii computed on the host
           from the GLOBAL decision vector; a failure exits non-zero.
``e2e``    the same through the host API: decision vector and multipliers in
           pinned host memory -> one H2D copy -> kernels -> one D2H copy of
           all five results; with more than one GPU through ONE solver-facing
           process (sharding.SolverFacingEvaluator): every rank moves its own
           slices of the global-order vectors in shared page-locked memory.
``e2e_solver_api``  (N = 1) the five IPOPT callbacks in IPOPT's order through
           nlp.GpuEvaluator.ipopt_eval with caller-owned PAGEABLE NumPy arrays
           for x, lambda, f, grad, g, Jacobian and Hessian values, a new x every
           step; ``e2e_solver_api_pinned``: the same arrays page-locked once
           (GpuEvaluator.pin), as a solver with stable buffers would.
``roofline``      HBM roofline of the fused per-sample kernel (a first pass of
           the same K steps with CUDA events around that kernel alone).
``cpu_baseline``  the CPU oracle (NumPy restatement of the reference's
           evaluation) timed on this box, rank 0, N = 1 only.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KIND = 'ml'
DIMS = (2, 1, 2)
N_PER_GPU = 1_000_000
DT = 0.05       # sample time of the continuous-time families (ZOH / noise discretisation)
METRIC = 'nlp_callback_evals_per_s'
UNIT = 'callback sets/s (f+grad+g+Jac+Hess, N=1e6 samples per set)'
FLUSH_BYTES = 256 << 20


SCALING = 'weak'
DEFAULT_WORKLOAD = True


def configure_workload(args):
    """Apply --kind / --dims / --n-per-gpu / --n-total to the module-level
    workload description."""
    global KIND, DIMS, N_PER_GPU, SCALING, UNIT, DEFAULT_WORKLOAD
    world = int(os.environ.get('WORLD_SIZE', 1))
    kind, dims = args.kind, tuple(int(c) for c in args.dims)
    n = args.n_per_gpu
    if args.n_total:
        if args.n_total % world:
            raise SystemExit('--n-total must be a multiple of the GPU count')
        n = args.n_total // world
        SCALING = 'strong'
        UNIT = ('callback sets/s (f+grad+g+Jac+Hess, '
                f'N={args.n_total} samples per set)')
    elif n != N_PER_GPU:
        UNIT = f'callback sets/s (f+grad+g+Jac+Hess, N={n} samples per set)'
    DEFAULT_WORKLOAD = (kind, dims, n, SCALING) == (KIND, DIMS, N_PER_GPU,
                                                    'weak')
    KIND, DIMS, N_PER_GPU = kind, dims, n


SCRIPT_OF = {'ml': 'attas_sp_ml', 'balanced': 'attas_sp_innov_bal / '
             'blackbox_innov_bal', 'ndisc_zoh': 'hfb320_sqrt_zoh / '
             'attas_sp_ml_ndisc', 'ml_zoh': 'attas_sp_ml_zoh',
             'innovation': 'attas_sp_innov', 'ml_balanced': 'mc_blackbox_cfem'}


def algorithmic_bytes_per_sample(nx, nu, ny):
    """every input read once, every output written
    once, per sample, for the per-sample functions."""
    nty = ny * (ny + 1) // 2
    jd = nx * (1 + 2 * nx + 2 * ny + nu)
    ji = 2 * nty + 2 * ny * nx + ny * nu + ny
    hd = nx * nx + nx * ny
    hi = nty + ny * nx
    ho = 2 * ny
    return 8 * ((nx + 2 * ny + nu) + 2 * (nx + ny) + ny + jd + ji
                + hd + hi + ho)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device),
                 f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [],
                    'note': 'nvidia-smi unavailable'}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap')
        for row in self.rows:
            parts = [p.strip() for p in row.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the GPU's NUMA node so that the pinned
    host buffers (first touch) and the DMA engines sit on the same socket.
    Returns (node or None, note)."""
    try:
        out = subprocess.run(
            ['nvidia-smi', '-i', str(device), '--query-gpu=pci.bus_id',
             '--format=csv,noheader'], capture_output=True, text=True,
            timeout=30).stdout.strip().lower()
        dom, rest = out.split(':', 1)
        bus = f'{dom[-4:]}:{rest}'
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as fh:
            node = int(fh.read())
        if node < 0:
            return None, f'{bus}: numa_node {node} (no NUMA information)'
        with open(f'/sys/devices/system/node/node{node}/cpulist') as fh:
            cpus = set()
            for part in fh.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node, f'{bus}: bound to {len(allowed)} CPUs of node {node}'
    except Exception as exc:
        return None, repr(exc)


class CudaArray:
    """Zero-copy torch view of a device buffer owned by a cfem handle."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {
            'shape': (n,), 'typestr': '<f8', 'data': (int(ptr), False),
            'version': 2}


def run_reference(args, out):
    """The reference's CPU evaluation (the NumPy oracle port: the reference's
    own stack -- ceacoest / sym2num -- is not installable), same workload."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from colloc_fem_code_b200 import families, synthetic
    from oracle import ref_models
    nx, nu, ny = DIMS
    N = N_PER_GPU
    exp = synthetic.experiment(0, N, nx, nu, ny)
    o = ref_models.make_problem(KIND, exp['y'], exp['u'], nx, dt=DT)
    p = families.make_problem(KIND, exp['y'], exp['u'], nx, dt=DT)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    o.constr_jac_ind()
    o.lag_hess_ind()

    def step(i):
        d = dvec + 1e-9 * i
        o.obj(d)
        o.obj_grad(d)
        o.constr(d)
        o.constr_jac_val(d)
        o.lag_hess_val(d, sigma, lam)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = f'full workload: N={N} samples per step, {args.steps} steps'
    out.emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': workload_config(1),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1,
                         'kind': 'port', 'sample': sample,
                         'host_cores_available': os.cpu_count()},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


def workload_config(world, reduce_mode='none'):
    nx, nu, ny = DIMS
    how = {'skip': 'NO reduction (debug only: results incomplete)',
           'peer': 'objective + parameter gradient: partial sums posted to '
                   'every peer inside the kernel over NVLink peer memory; '
                   'the rank-order sum of step k is finished inside the '
                   'kernel of step k+1 (by a collect kernel after the last '
                   'step), so ranks do not rendezvous at every step; no '
                   'NCCL call on the path',
           'peer_sync': 'objective + parameter gradient reduced inside the '
                        'kernel over NVLink peer memory, all ranks rendezvous '
                        'at the end of every kernel (no NCCL call on the '
                        'path)',
           'nccl': 'NCCL allreduce of objective + parameter gradient',
           'none': ''}[reduce_mode]
    return {
        'workload': (f'{SCRIPT_OF.get(KIND, KIND)}: family {KIND} '
                     f'(nx,nu,ny)=({nx},{nu},{ny}), synthetic trajectory, '
                     f'N={N_PER_GPU} samples per GPU'
                     if not DEFAULT_WORKLOAD else
                     'attas_sp_ml: MaximumLikelihoodDT (nx,nu,ny)=(2,1,2), '
                     f'synthetic trajectory, N={N_PER_GPU} samples per GPU'),
        'family': KIND, 'dims': list(DIMS),
        'n_samples_total': N_PER_GPU * world,
        'parallelism': 'single GPU' if world == 1 else
        f'time-sharded x{world}, 1-sample halo, {how}'
        + ('; `value` = this pipelined exchange, `value_sync` = every kernel '
           'waits for its own sums' if reduce_mode == 'peer' else ''),
        'l2': f'flushed between timed steps ({FLUSH_BYTES >> 20} MiB memset, '
              'outside the per-step events); working set per step '
              f'{algorithmic_bytes_per_sample(nx, nu, ny) * N_PER_GPU / 1e6:.0f}'
              ' MB > 126 MB L2',
    }


def cpu_baseline():
    """Oracle timed on one host core on a bounded sample of the workload."""
    from colloc_fem_code_b200 import families, synthetic
    from oracle import ref_models
    nx, nu, ny = DIMS
    N = N_PER_GPU
    exp = synthetic.experiment(0, N, nx, nu, ny)
    o = ref_models.make_problem(KIND, exp['y'], exp['u'], nx, dt=DT)
    p = families.make_problem(KIND, exp['y'], exp['u'], nx, dt=DT)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        o.obj(dvec)
        o.obj_grad(dvec)
        o.constr(dvec)
        o.constr_jac_val(dvec)
        o.lag_hess_val(dvec, sigma, lam)
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {'value': 1.0 / best, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': f'full workload N={N}, best of 3 callback sets after '
                      '1 warm-up (index arrays excluded)',
            'seconds_per_set': best,
            'host_cores_available': os.cpu_count()}


def run_ours(args, out):
    import torch
    import torch.distributed as dist
    from colloc_fem_code_b200 import backend, families, sharding, synthetic

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    numa_node, numa_note = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(
            'cuda', local_rank))

    nx, nu, ny = DIMS
    N = N_PER_GPU * world
    exp = synthetic.experiment(0, N, nx, nu, ny)
    problem = families.make_problem(KIND, exp['y'], exp['u'], nx, dt=DT)
    dvec, lam, sigma = synthetic.evaluation_point(problem, exp)
    ev = sharding.ShardedEvaluator(problem, rank, world, device=local_rank)
    h = ev.handle
    # one explicit stream for the kernels, the events and the NCCL calls
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.set_kernel_timing(True)
    ldvec = ev.shard.local_dvec(dvec)
    llam = ev.shard.local_multipliers(lam)
    del exp

    # ---- device-resident inputs (torch only owns the memory) -----------------
    d_dvec = torch.from_numpy(ldvec).cuda()
    d_lam = torch.from_numpy(llam).cuda()
    h.set_dvec_device(d_dvec.data_ptr())
    h.set_multipliers_device(sigma, d_lam.data_ptr())
    ptrs = h.device_ptrs()
    red = torch.as_tensor(CudaArray(ptrs['reduce'], ev.n_reduce),
                          device=f'cuda:{local_rank}')
    # cross-GPU reduction of [f, d f/d params]: fused into the kernel over
    # NVLink peer memory; CFEM_REDUCE=nccl selects all_reduce + a second kernel
    reduce_mode = 'none'
    if world > 1:
        reduce_mode = os.environ.get('CFEM_REDUCE', 'peer')
        if reduce_mode in ('peer', 'peer_sync'):
            # rank-local preparation, vote, then the rendezvous on all ranks
            if not ev.agree_on_peer_reduce(pipelined=reduce_mode == 'peer'):
                print(f'rank {rank}: peer reduce unavailable; using NCCL',
                      file=sys.stderr)
                reduce_mode = 'nccl'

    def device_step():
        h.set_dvec_device(d_dvec.data_ptr())      # "new x": invalidates
        h.eval(backend.ALL)
        if reduce_mode == 'nccl':
            dist.all_reduce(red)
            h.apply_reduced(ptrs['reduce'])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_pass(kernel_events):
        """W warm-up steps, then exactly K steps with events around every
        step (L2 flushed before each, outside the events)."""
        h.set_kernel_timing(kernel_events)
        starts = [torch.cuda.Event(enable_timing=True)
                  for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True)
                 for _ in range(args.steps)]
        for _ in range(args.warmup):
            h.flush_l2(FLUSH_BYTES)
            device_step()
        h.synchronize()
        sync_all()
        launches0 = h.launch_count
        t_wall = time.perf_counter()
        for i in range(args.steps):
            h.flush_l2(FLUSH_BYTES)
            starts[i].record()
            device_step()
            stops[i].record()        # no host synchronisation inside the loop
        host_enqueue = time.perf_counter() - t_wall
        h.synchronize()          # pipelined reduction: finishes the last step
        sync_all()
        out = {'wall': time.perf_counter() - t_wall,
               'host_enqueue': host_enqueue,
               'launches': h.launch_count - launches0,
               'step_ms': [s.elapsed_time(e) for s, e in zip(starts, stops)]}
        if kernel_events:
            out['kernel_ms'] = h.sample_kernel_ms_history(min(args.steps, 64))
        return out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)          # nvidia-smi start-up is over before timing
    # pass 1: CUDA events around the per-sample kernel itself (roofline); they
    # sit between the launches and cost a few microseconds per step, so the
    # throughput is taken from pass 2, the same K steps without them
    probe = timed_pass(True)
    timed = timed_pass(False)
    timed_sync = None
    if reduce_mode == 'peer':
        # the mode a sequential solver can use: post AND collect in one launch
        h.set_peer_mode(False)
        timed_sync = timed_pass(False)
    reduce_check = check_reduction(h, ev, problem, dvec, ldvec, world,
                                   local_rank, dist if world > 1 else None,
                                   torch)
    if reduce_mode == 'peer':
        h.set_peer_mode(True)
    kernel_ms = probe['kernel_ms']
    wall, host_enqueue = timed['wall'], timed['host_enqueue']
    launches, step_ms = timed['launches'], timed['step_ms']
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64,
                            device=f'cuda:{local_rank}')
    mine = torch.tensor([sum(step_ms) / args.steps, float(np.mean(kernel_ms))],
                        dtype=torch.float64, device=f'cuda:{local_rank}')
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    total_ms = float(total_ms.item())
    sync_ms = None
    if timed_sync is not None:
        t = torch.tensor([sum(timed_sync['step_ms'])], dtype=torch.float64,
                         device=f'cuda:{local_rank}')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sync_ms = float(t.item())
    per_rank = {'ms_per_step': [round(float(t[0]), 5) for t in per_rank],
                'kernel_ms': [round(float(t[1]), 5) for t in per_rank]}

    # ---- end to end through the host API (pinned host buffers) ---------------
    e2e_steps = 30              # >= 30: min and median are reported too
    e2e_step_s = []
    e2e_ms, e2e_h2d, e2e_d2h, solver_api = 0.0, 0, 0, None
    e2e_note = 'skipped (--no-e2e)'
    if not args.no_e2e:
        (e2e_ms, e2e_h2d, e2e_d2h, e2e_note, solver_api) = run_e2e(
            args, torch, dist, backend, sharding, problem, ev, h, rank,
            world, local_rank, reduce_mode, red, ptrs, dvec, lam, sigma,
            ldvec, llam, e2e_steps, e2e_step_s, sync_all)
    else:
        e2e_step_s.append(0.0)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # weak: N_PER_GPU-sample callback sets; strong: whole-trajectory sets
        sets = args.steps * (world if SCALING == 'weak' else 1)
        e2e_sets = e2e_steps * (world if SCALING == 'weak' else 1)
        value = sets / (total_ms * 1e-3)
        balg = algorithmic_bytes_per_sample(nx, nu, ny)
        kms = float(np.mean(kernel_ms))
        peak, peak_kind = measured_peak()
        achieved = balg * ev.shard.n_local / (kms * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': SCALING, 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': workload_config(world, reduce_mode),
            'samples_per_s': value * N_PER_GPU
            * (world if SCALING == 'strong' else 1),
            'numa_node': numa_node, 'numa_note': numa_note,
            'per_rank': per_rank,
            'ms_per_step_with_kernel_events': sum(probe['step_ms'])
            / args.steps,
            'gpu_launches': int(launches),
            'wall_s_timed_region': wall,
            'host_enqueue_s': host_enqueue,
            'clocks': clocks,
            'e2e': {
                'value': e2e_sets / (e2e_ms * 1e-3) if e2e_ms else None,
                'unit': UNIT,
                'h2d_bytes_per_step': e2e_h2d, 'd2h_bytes_per_step': e2e_d2h,
                'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps,
                'ms_per_step_min': 1e3 * min(e2e_step_s),
                'ms_per_step_median': 1e3 * float(np.median(e2e_step_s)),
                'note': e2e_note},
            'reduce_check': reduce_check,
            'roofline': {
                'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(),
                'kernel': 'cfem_sample_kernel_m31',
                'kernel_ms': kms, 'algorithmic_bytes_per_sample': balg,
                'samples_per_launch': ev.shard.n_local,
                'peak_source': f'{peak_kind} (MEASURED_PEAKS.json hbm_gbs, '
                               'burst copy)',
                'frac_of_nominal_8TBs': achieved / 8000.0},
        }
        if sync_ms is not None:
            line['value_sync'] = sets / (sync_ms * 1e-3)
            line['ms_per_step_sync'] = sync_ms / args.steps
        if solver_api is not None:
            line['e2e_solver_api'] = solver_api['pageable']
            line['e2e_solver_api_pinned'] = solver_api['pinned']
        if world == 1 and not args.no_cpu_baseline and DEFAULT_WORKLOAD:
            line['cpu_baseline'] = cpu_baseline()
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not reduce_check['ok']:
        raise SystemExit(f'reduce_check failed: {reduce_check}')


def run_e2e(args, torch, dist, backend, sharding, problem, ev, h, rank, world,
            local_rank, reduce_mode, red, ptrs, dvec, lam, sigma, ldvec, llam,
            e2e_steps, e2e_step_s, sync_all):
    """The end-to-end legs: host buffers -> H2D -> kernels -> D2H, every
    step (see the module docstring)."""
    sfe = None
    if world > 1:
        # ONE solver-facing process (rank 0) in front of all shards: x, lambda
        # and the five results live in shared page-locked host vectors in the
        # global (IPOPT) order; every rank DMAs only its own pieces over its
        # own PCIe link (sharding.SolverFacingEvaluator).
        hook = None
        if reduce_mode == 'nccl':
            def hook(handle):
                dist.all_reduce(red)
                handle.apply_reduced(ptrs['reduce'])
        try:
            sfe = sharding.SolverFacingEvaluator(
                problem, ev.shard, h, rank, world,
                broadcast=lambda box: dist.broadcast_object_list(box, src=0),
                barrier=dist.barrier, reduce_hook=hook, lib=ev.lib)
        except (OSError, RuntimeError, MemoryError) as exc:
            print(f'rank {rank}: shared host vectors unavailable ({exc!r}); '
                  'per-rank end-to-end instead', file=sys.stderr)
        agreed = torch.tensor([sfe is not None], dtype=torch.int32,
                              device=f'cuda:{local_rank}')
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if sfe is not None and int(agreed.item()) == 0:
            sfe.sv.close()
            sfe = None
    if sfe is None:
        host = backend.HostBuffers(h)
        host.dvec[:] = ldvec
        host.lam[:] = llam

        def e2e_step(i):
            host.dvec[0] = ldvec[0] + 1e-12 * i      # a new x every step
            if world == 1:
                # x up, f|grad|g|Jacobian kernels and values down; lambda up
                # and the Hessian group beside them (cfem_eval_callback_set)
                host.callback_set(sigma)
                return
            host.upload(sigma)           # one H2D copy: [x | lambda]
            h.eval(backend.ALL)
            if reduce_mode == 'nccl':
                dist.all_reduce(red)
                h.apply_reduced(ptrs['reduce'])
            host.fetch_all()

        for i in range(2):
            e2e_step(i)
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(e2e_steps):
            t_step = time.perf_counter()
            e2e_step(2 + i)             # ends with a stream synchronise
            e2e_step_s.append(time.perf_counter() - t_step)
        e1.record()
        sync_all()
        e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64,
                              device=f'cuda:{local_rank}')
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2e_ms.item())
        e2e_h2d = int(8 * (h.ndec + h.ncons))
        e2e_d2h = int(8 * (1 + h.ndec + h.ncons + h.nnz_jac + h.nnz_hess))
        e2e_note = ('pinned host [dvec | lambda] -> H2D -> fused kernels -> '
                    'D2H of [f | grad | g | Jacobian | Hessian values] (per '
                    'rank), one C-ABI call per set '
                    '(cfem_eval_callback_set: lambda goes up while the first '
                    'results come down); CUDA events' if world == 1 else
                    'pinned host [dvec | lambda] -> one H2D copy -> fused '
                    'kernels -> one D2H copy of [f | grad | g | Jacobian | '
                    'Hessian values] (per rank); CUDA events')
    else:
        e2e_ms = 0.0
        if rank == 0:
            sv = sfe.sv
            sv.dvec[:] = dvec
            sv.lam[:] = lam
            sv.sigma = sigma
            request = sfe.X | sfe.LAMBDA | backend.ALL
            for i in range(2):
                sv.dvec[0] = dvec[0] + 1e-12 * i
                sfe._request(request)
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                t_step = time.perf_counter()
                sv.dvec[0] = dvec[0] + 1e-12 * (2 + i)
                sfe._request(request)
                e2e_step_s.append(time.perf_counter() - t_step)
            e2e_ms = 1e3 * (time.perf_counter() - t0)
            sfe.stop()
        else:
            sfe.serve()
        sync_all()
        g = ev.shard.glob
        # parameters / parameter-only multipliers are uploaded by every rank
        e2e_h2d = int(8 * sum(int(pc[:, 2].sum()) for pc in sfe._in.values()))
        e2e_d2h = int(8 * sum(int(pc[:, 2].sum()) for pc in sfe._out.values()))
        e2e_note = ('one solver-facing process: x, lambda and results in '
                    'shared page-locked host vectors in the global order '
                    f'(N={g.N} samples); every rank moves its own pieces '
                    '(H2D -> fused kernels + cross-GPU reduction -> D2H); '
                    'host wall clock on rank 0 around complete requests; '
                    'bytes are per rank; page-locked: ' + str(sfe.pinned))
        sfe.close()
    solver_api = None
    if world == 1:
        res = {k: np.array(getattr(host, k)) for k in
               ('f', 'grad', 'g', 'jac', 'hess')}
        host.close()
        solver_api = {
            'pageable': solver_api_leg(problem, dvec, lam, sigma, e2e_steps,
                                       False, res, 2 + e2e_steps - 1),
            'pinned': solver_api_leg(problem, dvec, lam, sigma, e2e_steps,
                                     True, res, 2 + e2e_steps - 1)}
    return e2e_ms, e2e_h2d, e2e_d2h, e2e_note, solver_api


def check_reduction(h, ev, problem, dvec, ldvec, world, device, dist, torch):
    """Every rank: the reduced objective and the sRp-diagonal gradient entries
    it holds after the timed passes against the closed form computed on the
    host from the GLOBAL decision vector (symfem.py:61-65, adfem.py:119), and
    its en-block of the gradient (= -en, exact).  MAX / MIN over ranks."""
    from colloc_fem_code_b200 import backend, models
    h.synchronize()
    f = float(h.fetch(backend.F)[0])
    grad = h.fetch(backend.GRAD)
    var = problem.variables(dvec)
    en = var['en']
    ny = en.shape[1]
    diag = var['sRp_tril'][models.tril_diag(ny)]
    n_total = en.shape[0]
    f_ref = -0.5 * float(np.sum(en * en)) - n_total * float(np.log(diag).sum())
    off = dict(zip(problem.structure.var_names, h.layout()['var_offset']))
    lo = int(off['sRp_tril'])
    g_srp = grad[lo:lo + var['sRp_tril'].size][models.tril_diag(ny)]
    rel = max(abs(f - f_ref) / abs(f_ref),
              float(np.max(np.abs(g_srp + n_total / diag) * diag / n_total)))
    eo = int(off['en'])
    n_en = ev.shard.n_local * ny
    en_exact = bool(np.array_equal(grad[eo:eo + n_en],
                                   -ldvec[eo:eo + n_en]))
    ok = rel <= 1e-12 and en_exact
    if dist is not None:
        t = torch.tensor([rel, 0.0 if ok else 1.0, f], dtype=torch.float64,
                         device=f'cuda:{device}')
        gathered = [t.clone() for _ in range(world)]
        dist.all_gather(gathered, t)
        rel = max(float(g[0]) for g in gathered)
        same_bits = all(float(g[2]) == float(gathered[0][2])
                        for g in gathered)
        ok = all(float(g[1]) == 0.0 for g in gathered) and same_bits
    else:
        same_bits = True
    return {'ok': bool(ok), 'rel_err': rel, 'tol': 1e-12, 'f': f,
            'f_closed_form': f_ref, 'grad_en_block_exact': en_exact,
            'same_bits_on_all_ranks': same_bits, 'ranks': world,
            'what': 'f and d f/d sRp_ii vs -1/2 sum en^2 - N sum log sRp_ii '
                    'and -N/sRp_ii from the global dvec, every rank'}


def solver_api_leg(problem, dvec, lam, sigma, steps, pinned, expect, last_i):
    """IPOPT's callback sequence (eval_f, eval_grad_f, eval_g, eval_jac_g,
    eval_h -- IpStdCInterface.h) through nlp.GpuEvaluator.ipopt_eval with
    caller-owned NumPy arrays and a new x every step; wall clock (every
    callback returns with its result in the caller's array)."""
    from colloc_fem_code_b200 import backend, nlp
    ev = nlp.GpuEvaluator(problem)
    n, m = problem.ndec, problem.ncons
    x, lam_a = np.array(dvec), np.array(lam)
    out = {'f': np.empty(1), 'grad': np.empty(n), 'g': np.empty(m),
           'jac': np.empty(problem.nnzjac), 'hess': np.empty(problem.nnzhess)}
    if pinned:
        ev.pin(x, lam_a, *[out[k] for k in ('grad', 'g', 'jac', 'hess')])

    def step(i):
        x[0] = dvec[0] + 1e-12 * i          # a new x every step
        ev.ipopt_eval(backend.F, x, True, out['f'])
        ev.ipopt_eval(backend.GRAD, x, False, out['grad'])
        ev.ipopt_eval(backend.G, x, False, out['g'])
        ev.ipopt_eval(backend.JAC, x, False, out['jac'])
        ev.ipopt_eval(backend.HESS, x, False, out['hess'], sigma, lam_a)

    for i in range(2):
        step(i)
    times = []
    t0 = time.perf_counter()
    for i in range(steps):
        t1 = time.perf_counter()
        step(last_i - steps + 1 + i)
        times.append(time.perf_counter() - t1)
    total = time.perf_counter() - t0
    # the last step used the same x as the last step of the C-ABI leg
    for k in ('f', 'grad', 'g', 'jac', 'hess'):
        np.testing.assert_allclose(out[k], expect[k], rtol=1e-12, atol=1e-300,
                                   err_msg=f'e2e_solver_api: {k}')
    groups = ev.kernel_groups
    ev.close()
    return {'value': steps / total, 'unit': UNIT,
            'h2d_bytes_per_step': int(8 * (n + m)),
            'd2h_bytes_per_step': int(8 * (1 + n + m + problem.nnzjac
                                           + problem.nnzhess)),
            'ms_per_step': 1e3 * total / steps, 'steps': steps,
            'ms_per_step_min': 1e3 * min(times),
            'ms_per_step_median': 1e3 * float(np.median(times)),
            'kernel_groups_per_step': groups / (steps + 2),
            'host_arrays': 'page-locked once (GpuEvaluator.pin)' if pinned
            else 'pageable NumPy (threaded staging inside the C ABI)',
            'checked_against_c_abi_leg': True}


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/), or None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path) as fh:
            return json.load(fh)['cfem_sample_kernel_m31']['dram_bytes']
    except Exception:
        return None


class OneLineStdout:
    """Everything libraries print to fd 1 during the run (e.g. NCCL's version
    banner) goes to stderr; only the final JSON line reaches stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.write(self._saved, (line + '\n').encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    # other script shapes / strong scaling (profiles/ tables); the default
    # line -- attas_sp_ml, (2,1,2), 1e6 samples per GPU, weak -- is untouched
    ap.add_argument('--kind', default=KIND,
                    help='problem family (families.make_problem)')
    ap.add_argument('--dims', default=''.join(str(d) for d in DIMS),
                    help='nx nu ny as three digits, e.g. 533')
    ap.add_argument('--n-per-gpu', type=int, default=N_PER_GPU)
    ap.add_argument('--n-total', type=int, default=0,
                    help='strong scaling: fixed trajectory length split over '
                         'the GPUs (overrides --n-per-gpu)')
    ap.add_argument('--no-e2e', action='store_true',
                    help='device-resident numbers only (table runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    configure_workload(args)
    with OneLineStdout() as out:
        if args.impl == 'reference':
            run_reference(args, out)
        else:
            run_ours(args, out)


if __name__ == '__main__':
    main()
