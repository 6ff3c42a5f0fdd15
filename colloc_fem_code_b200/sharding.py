"""Time-sharding of one long trajectory across GPUs (one process per GPU).

The only coupling along time in the reference's problems is nearest-neighbour:
row k of ``dynamics`` reads ``x[k]`` and ``x[k+1]`` (the shifted views
``xprev`` / ``xnext`` of /root/reference/fem.py:47-52).  A contiguous block of
samples ``[k0, k1)`` per rank is therefore exact with a ONE-sample halo of the
variables that are read one step ahead; every rank owns a disjoint slice of the
constraint vector, of the per-sample part of the gradient and of every
Jacobian / Hessian value block.  What has to cross ranks per callback is the
scalar objective and the parameter entries of the gradient (the sums over
samples of /root/reference/adfem.py:119): one all-reduce of ``n_reduce``
doubles (NCCL over NVLink on the GPU box, gloo in the CPU tests).

This module is the host-side bookkeeping: which rows of the global decision /
multiplier / result vectors a rank owns, expressed with the N-independent
``optim.Structure`` so that it mirrors ``compute_layout()`` of
``csrc/cfem_host.inl``.  The parameter-only constraints are evaluated by every
rank redundantly (they are O(1)); rank ``world-1`` (the one without halo) is
their owner when results are gathered.
"""

import os

import numpy as np


def split_samples(N, world):
    """Contiguous, balanced sample ranges ``[(k0, k1)] * world``."""
    base, rem = divmod(N, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class Layout:
    """Offsets of a problem with ``N`` samples (``halo`` as in cfem_create)."""

    def __init__(self, structure, N, halo=0):
        st = structure
        self.N, self.halo = N, halo
        self.var_rows = st.var_rows(N, halo)
        self.fun_rows = st.fun_rows(N, halo)
        off = 0
        self.var_off = []
        for v, rows in zip(st.vars, self.var_rows):
            self.var_off.append(off)
            off += rows * v['core']
        self.ndec = off
        off = 0
        self.cons_off = {}
        for fi, f in enumerate(st.funs):
            if f['cons_index'] >= 0:
                self.cons_off[fi] = off
                off += self.fun_rows[fi] * f['out_core']
        self.ncons = off
        off = 0
        self.jac_off = []
        for b in st.jac_blocks:
            self.jac_off.append(off)
            off += self.fun_rows[b['fun']] * b['c']
        self.nnz_jac = off
        off = 0
        self.hess_off = []
        for b in st.hess_blocks:
            self.hess_off.append(off)
            off += self.fun_rows[b['fun']] * b['c']
        self.nnz_hess = off


class TimeShard:
    """Rank-local view of a time-sharded problem."""

    def __init__(self, structure, N, rank, world):
        if N < 2 * world:
            raise ValueError('need at least two samples per shard')
        self.st = structure
        self.N, self.rank, self.world = N, rank, world
        self.k0, self.k1 = split_samples(N, world)[rank]
        self.n_local = self.k1 - self.k0
        self.halo = 0 if rank == world - 1 else 1
        self.glob = Layout(structure, N, 0)
        self.loc = Layout(structure, self.n_local, self.halo)
        self.owns_params = rank == world - 1

    # -- data ------------------------------------------------------------------
    def local_data(self):
        """Rank-local slices of the per-sample data arrays (y, u)."""
        out = []
        for d in self.st.data:
            rows = self.n_local + d['r0'] + self.halo * d['hshift']
            out.append(np.ascontiguousarray(
                d['source'][self.k0:self.k0 + rows]))
        return out

    # -- inputs ------------------------------------------------------------------
    def input_pieces(self, kind):
        """``(local offset, global offset, length)`` rows (in doubles) of the
        rank-local decision vector (``'dvec'``) / multipliers (``'lam'``): the
        argument of ``cfem_upload_pieces``.  Parameters (and the multipliers
        of the parameter-only constraints) are replicated on every rank; the
        halo row of a per-sample variable is read from the right neighbour's
        range of the global vector."""
        rows = []
        if kind == 'dvec':
            for i, v in enumerate(self.st.vars):
                c = v['core']
                g0 = self.glob.var_off[i] + (self.k0 * c if v['per_sample']
                                             else 0)
                rows.append((self.loc.var_off[i], g0,
                             self.loc.var_rows[i] * c))
        elif kind == 'lam':
            for fi, f in enumerate(self.st.funs):
                if f['cons_index'] < 0:
                    continue
                c = f['out_core']
                g0 = self.glob.cons_off[fi] + (self.k0 * c if f['per_sample']
                                               else 0)
                rows.append((self.loc.cons_off[fi], g0,
                             self.loc.fun_rows[fi] * c))
        else:
            raise KeyError(kind)
        return np.asarray(rows, dtype=np.int64).reshape(-1, 3)

    def result_pieces(self, kind):
        """``(local offset, global offset, length)`` rows of the part of a
        result this rank owns (``cfem_fetch_pieces``); ``'f'`` belongs to the
        rank that owns the parameter-only functions."""
        if kind == 'f':
            rows = [(0, 0, 1)] if self.owns_params else []
        else:
            rows = [(l0, g0, n) for g0, l0, n in self._pairs(kind)]
        return np.asarray(rows, dtype=np.int64).reshape(-1, 3)

    def _gather(self, kind, vec, size):
        out = np.empty(size)
        for l0, g0, n in self.input_pieces(kind):
            out[l0:l0 + n] = vec[g0:g0 + n]
        return out

    def local_dvec(self, dvec):
        return self._gather('dvec', dvec, self.loc.ndec)

    def local_multipliers(self, lam):
        return self._gather('lam', lam, self.loc.ncons)

    # -- results: (global slice, local slice) pairs ------------------------------
    def _pairs(self, kind):
        st = self.st
        pairs = []
        if kind == 'g':
            for fi, f in enumerate(st.funs):
                if f['cons_index'] < 0:
                    continue
                if not f['per_sample'] and not self.owns_params:
                    continue
                c = f['out_core']
                n = self.loc.fun_rows[fi] * c
                g0 = self.glob.cons_off[fi] + (self.k0 * c if f['per_sample']
                                               else 0)
                pairs.append((g0, self.loc.cons_off[fi], n))
        elif kind in ('jac', 'hess'):
            blocks = st.jac_blocks if kind == 'jac' else st.hess_blocks
            goff = self.glob.jac_off if kind == 'jac' else self.glob.hess_off
            loff = self.loc.jac_off if kind == 'jac' else self.loc.hess_off
            for bi, b in enumerate(blocks):
                f = st.funs[b['fun']]
                if not f['per_sample'] and not self.owns_params:
                    continue
                n = self.loc.fun_rows[b['fun']] * b['c']
                g0 = goff[bi] + (self.k0 * b['c'] if f['per_sample'] else 0)
                pairs.append((g0, loff[bi], n))
        elif kind == 'grad':
            for i, v in enumerate(st.vars):
                c = v['core']
                if v['per_sample']:
                    # own rows only: the halo row belongs to the right neighbour
                    rows = self.n_local + (v['r0'] if self.halo == 0 else 0)
                    pairs.append((self.glob.var_off[i] + self.k0 * c,
                                  self.loc.var_off[i], rows * c))
                elif self.owns_params:
                    pairs.append((self.glob.var_off[i], self.loc.var_off[i],
                                  self.loc.var_rows[i] * c))
        else:
            raise KeyError(kind)
        return pairs

    def scatter(self, kind, local, out):
        """Write this rank's part of a result into the global vector."""
        for g0, l0, n in self._pairs(kind):
            out[g0:g0 + n] = local[l0:l0 + n]
        return out

    def global_size(self, kind):
        return {'g': self.glob.ncons, 'jac': self.glob.nnz_jac,
                'hess': self.glob.nnz_hess, 'grad': self.glob.ndec}[kind]


class ShardedEvaluator:
    """One rank of a time-sharded evaluation.

    ``allreduce(array_or_tensor)`` sums the ``[n_reduce]`` vector (objective,
    parameter-gradient entries) over ranks in place; ``evaluate`` returns this
    rank's local results with the all-reduced objective / parameter gradient
    written back on the device (``cfem_apply_reduced``).
    """

    def __init__(self, problem, rank, world, device=0):
        from . import backend
        self.problem = problem
        st = problem.structure
        self.shard = TimeShard(st, st.N, rank, world)
        self.lib = backend.Library.for_structure(st)
        self.handle = backend.Handle(self.lib, self.shard.n_local,
                                     self.shard.local_data(),
                                     st.scalar_values, halo=self.shard.halo,
                                     device=device)
        h, loc = self.handle, self.shard.loc
        got = (h.ndec, h.ncons, h.nnz_jac, h.nnz_hess)
        want = (loc.ndec, loc.ncons, loc.nnz_jac, loc.nnz_hess)
        if got != want:
            raise backend.CfemError(f'shard layout mismatch: {got} vs {want}')
        self.n_reduce = len(self.lib.model['reduce'])

    def prepare_peer_reduce(self, group=None):
        """Rank-local half of :meth:`enable_peer_reduce`: allocate and zero the
        symmetric inbox.  No collective is issued, so a failure here (no
        symmetric memory, no P2P) can be voted on before any rank enters the
        rendezvous (:func:`agree_on_peer_reduce`)."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        world = dist.get_world_size(group)
        n_inbox, n_flag = self.handle.peer_layout(world)
        buf = symm.empty(n_inbox + n_flag, dtype=torch.float64,
                         device=torch.device('cuda', torch.cuda.current_device()))
        buf.zero_()
        self._peer_buf = (buf, n_inbox)

    def enable_peer_reduce(self, group=None, pipelined=False):
        """Switch from ``all_reduce`` + ``cfem_apply_reduced`` to the fused
        in-kernel exchange over NVLink peer memory (``cfem_set_peers``).

        PyTorch is only the plumbing here: a symmetric-memory allocation
        (``torch.distributed._symmetric_memory``) gives every rank a device
        pointer to every other rank's inbox; the exchange itself is done by
        the last CTA of the per-sample kernel.  ``pipelined``: the kernel only
        posts its partial sums; a collect kernel on a side stream finishes the
        sum beside the next launch (``cfem_set_peer_mode``).
        """
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if getattr(self, '_peer_buf', None) is None:
            self.prepare_peer_reduce(group)
        buf, n_inbox = self._peer_buf
        hdl = symm.rendezvous(buf, group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        ptrs[rank] = buf.data_ptr()     # own inbox: the local mapping
        torch.cuda.synchronize()
        dist.barrier(group)             # every inbox is zeroed before any store
        self.handle.set_peers(rank, world, ptrs,
                              [p + 8 * n_inbox for p in ptrs])
        self.handle.set_peer_mode(pipelined)
        self._peer = (buf, hdl)         # keep the mapping alive
        return True

    def agree_on_peer_reduce(self, group=None, pipelined=False):
        """Collective: every rank tries the rank-local preparation, the ranks
        vote, and only if ALL succeeded do they enter the rendezvous -- the
        sequence of collectives is the same on every rank whatever fails
        where.  Returns True if the fused exchange is on."""
        import torch
        import torch.distributed as dist
        try:
            self.prepare_peer_reduce(group)
            ok = 1
        except Exception:
            self._peer_buf = None
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32,
                            device=torch.device('cuda',
                                                torch.cuda.current_device()))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self._peer_buf = None
            return False
        self.enable_peer_reduce(group, pipelined)
        return True

    def set_point(self, dvec, obj_factor, lam):
        self.handle.set_dvec(self.shard.local_dvec(dvec))
        self.handle.set_multipliers(obj_factor,
                                    self.shard.local_multipliers(lam))


# -----------------------------------------------------------------------------
# one solver process in front of all ranks
# -----------------------------------------------------------------------------

class SharedVectors:
    """The solver-facing vectors of a time-sharded problem -- decision vector,
    multipliers, objective, gradient, constraints, Jacobian and Hessian values,
    all in the GLOBAL (IPOPT) order -- in ONE shared-memory segment that every
    rank maps.  Each rank page-locks its mapping (``cfem_host_register``) so
    that its GPU's DMA engines move the rank's pieces straight between this
    memory and the device arrays over the rank's own PCIe link; the solver
    process reads complete vectors without any gather step.

    The segment starts with a control block of int64 words:
    ``[0]`` request sequence number, ``[1]`` request (result mask | CFEM_X |
    CFEM_LAMBDA, or ``STOP``), ``[2]`` obj_factor (float64 bits),
    ``[8 + r]`` sequence number rank ``r`` has completed.
    """

    FIELDS = ('dvec', 'lam', 'f', 'grad', 'g', 'jac', 'hess')
    STOP = -1
    _CTRL = 8

    def __init__(self, sizes, world, path=None):
        import mmap
        self.world = world
        self.sizes = {k: int(sizes[k]) for k in self.FIELDS}
        nctrl = self._CTRL + world
        nctrl += (-nctrl) % 8                   # 64-byte aligned vectors
        total = 8 * (nctrl + sum(n + (-n) % 8 for n in self.sizes.values()))
        self.created = path is None
        self._fd = None
        if path is None:
            path, fd = self._create(total)
        else:
            fd = os.open(path, os.O_RDWR)
        self.path = path
        try:
            self._map = mmap.mmap(fd, total)
        finally:
            if fd != self._fd:
                os.close(fd)
        self.ctrl = np.frombuffer(self._map, dtype=np.int64, count=nctrl)
        self._sigma = np.frombuffer(self._map, dtype=np.float64, count=nctrl)
        off = 8 * nctrl
        for k in self.FIELDS:
            n = self.sizes[k]
            setattr(self, k, np.frombuffer(self._map, dtype=np.float64,
                                           count=n, offset=off))
            off += 8 * (n + (-n) % 8)
        self.nbytes = total
        self._registered = None

    def _create(self, total):
        """Backing store: a tmpfs file under /dev/shm if it has room (checked
        up front -- a too-small /dev/shm would fail later with SIGBUS); else an
        anonymous memfd that the other ranks open through ``/proc/<pid>/fd``.
        Pages are NOT allocated here: every rank first-touches its own pieces
        (``SolverFacingEvaluator``), which places them on the NUMA node of the
        rank that will DMA them."""
        import uuid
        path = f'/dev/shm/cfem_{os.getpid()}_{uuid.uuid4().hex[:12]}'
        try:
            vfs = os.statvfs('/dev/shm')
            if vfs.f_bavail * vfs.f_frsize > total + (64 << 20):
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                try:
                    os.ftruncate(fd, total)
                    return path, fd
                except OSError:
                    os.close(fd)
                    os.unlink(path)
        except OSError:
            pass
        fd = os.memfd_create('cfem_shared_vectors')
        os.ftruncate(fd, total)
        self._fd = fd                   # stays open until unlink()
        return f'/proc/{os.getpid()}/fd/{fd}', fd

    def unlink(self):
        """Remove the name (the mappings stay valid); creator only."""
        if self.created and self.path:
            if self._fd is not None:
                os.close(self._fd)
                self._fd = None
            else:
                try:
                    os.unlink(self.path)
                except FileNotFoundError:
                    pass
            self.path = None

    def page_lock(self, lib):
        base = self.ctrl.ctypes.data
        if lib.cfem_host_register(base, self.nbytes) == 0:
            self._registered = (lib, base)
        return self._registered is not None

    def close(self):
        if self._registered:
            lib, base = self._registered
            lib.cfem_host_unregister(base)
            self._registered = None
        self.unlink()

    @property
    def sigma(self):
        return float(self._sigma[2])

    @sigma.setter
    def sigma(self, value):
        self._sigma[2] = value


class SolverFacingEvaluator:
    """All ranks of a time-sharded problem behind ONE NLP solver process.

    Rank 0 runs the solver (``nlp.make_solver(problem, ..., evaluator=this)``
    -- this object has the ``nlp.Evaluator`` surface); the other ranks call
    ``serve()``.  A callback posts a request in the control block of the
    ``SharedVectors``; every rank (rank 0 included) then uploads ITS pieces of
    x / lambda, runs the fused kernels on its shard (objective and parameter
    gradient are reduced across GPUs inside the kernel over NVLink peer memory,
    or by ``allreduce`` when that is unavailable), writes ITS pieces of the
    results into the shared vectors and acknowledges.  No rank ever touches
    another rank's slice and nothing is gathered through the solver process.

    ``handle`` needs ``upload_pieces``, ``fetch_pieces``, ``set_obj_factor``,
    ``eval`` and ``synchronize`` (``backend.Handle``); ``reduce_hook(handle)``,
    if given, is called after every evaluation that produced the objective or
    the gradient (e.g. ``all_reduce`` + ``apply_reduced``).  ``broadcast`` is
    ``torch.distributed.broadcast_object_list``-like and only used once, to
    share the segment's name.
    """

    RESULTS = ((1, 'f'), (2, 'grad'), (4, 'g'), (8, 'jac'), (16, 'hess'))
    X, LAMBDA = 32, 64

    def __init__(self, problem, shard, handle, rank, world, broadcast,
                 barrier, reduce_hook=None, lib=None):
        self.problem, self.shard, self.h = problem, shard, handle
        self.rank, self.world = rank, world
        self.reduce_hook = reduce_hook
        self.n, self.m = shard.glob.ndec, shard.glob.ncons
        sizes = {'dvec': self.n, 'lam': self.m, 'f': 1, 'grad': self.n,
                 'g': self.m, 'jac': shard.glob.nnz_jac,
                 'hess': shard.glob.nnz_hess}
        if rank == 0:
            try:
                self.sv = SharedVectors(sizes, world)
            except Exception:
                broadcast([None])       # the other ranks must not hang
                raise
            self.sv.ctrl[:] = 0
            broadcast([self.sv.path])
        else:
            box = [None]
            broadcast(box)
            if box[0] is None:
                raise RuntimeError('rank 0 could not create the shared '
                                   'host vectors')
            self.sv = SharedVectors(sizes, world, path=box[0])
        barrier()                       # everybody has mapped the segment
        self.sv.unlink()
        self._in = {k: shard.input_pieces(k) for k in ('dvec', 'lam')}
        self._out = {k: shard.result_pieces(k)
                     for _, k in self.RESULTS}
        # first touch: the pages of a rank's pieces are allocated by that rank
        # (its NUMA node, if the process is bound to the GPU's node) before
        # they are page-locked; fresh shared-memory pages are zero-filled
        for name, pieces in list(self._out.items()) + list(self._in.items()):
            vec = getattr(self.sv, name)
            for _, g0, n in pieces:
                if name in ('dvec', 'lam') and n < 4096:
                    continue            # replicated parameters: whoever is first
                vec[g0:g0 + n] = 0.0
        barrier()
        self._lib = lib
        self.pinned = self.sv.page_lock(lib) if lib is not None else False
        self._seq = 0
        self._have_x = False
        self._fresh = 0
        self.seconds = 0.0
        self.calls = 0
        self.kernel_groups = 0

    # -- what every rank does for one request ----------------------------------
    def _execute(self, request):
        h, sv = self.h, self.sv
        if request & self.X:
            h.upload_pieces(self.X, sv.dvec, self._in['dvec'])
        if request & self.LAMBDA:
            h.upload_pieces(self.LAMBDA, sv.lam, self._in['lam'])
            h.set_obj_factor(sv.sigma)
        mask = request & 31
        h.eval(mask)
        if self.reduce_hook is not None and mask & 3:
            self.reduce_hook(h)
        for bit, name in self.RESULTS:
            if mask & bit and len(self._out[name]):
                h.fetch_pieces(bit, getattr(sv, name), self._out[name])
        h.synchronize()

    TIMEOUT_S = 600.0       # a rank that died must not hang the others
    #: ranks > 0 wait for the solver between callbacks (its KKT factorisation
    #: can take long), but not for ever: rank 0 may have died without STOP
    SERVE_TIMEOUT_S = float(os.environ.get('CFEM_SERVE_TIMEOUT_S', 6 * 3600))

    def _publish(self, index, value):
        """Store a control word AFTER everything written before it (vectors,
        request, obj_factor): release store through the C ABI where the
        library provides it (needed on weakly ordered hosts, e.g. aarch64);
        a plain store is enough on x86."""
        store = getattr(self._lib, 'cfem_store_release_i64', None) \
            if self._lib is not None else None
        if store is not None:
            store(self.sv.ctrl.ctypes.data + 8 * index, int(value))
        else:
            self.sv.ctrl[index] = value

    def _observe(self, index):
        load = getattr(self._lib, 'cfem_load_acquire_i64', None) \
            if self._lib is not None else None
        if load is not None:
            return int(load(self.sv.ctrl.ctypes.data + 8 * index))
        return int(self.sv.ctrl[index])

    @classmethod
    def _wait(cls, cond, timeout=None):
        import time
        t0 = time.perf_counter()
        while not cond():
            waited = time.perf_counter() - t0
            if waited > 2e-3:
                time.sleep(1e-4)        # long host phases: stop burning a core
            if timeout is not None and waited > timeout:
                raise TimeoutError('time-sharded evaluation: a rank did not '
                                   f'answer within {timeout:.0f} s')

    def serve(self):
        """Ranks other than 0: execute requests until the solver stops."""
        ctrl = self.sv.ctrl
        seq = 0
        while True:
            self._wait(lambda: self._observe(0) != seq, self.SERVE_TIMEOUT_S)
            seq = self._observe(0)
            request = int(ctrl[1])
            if request == SharedVectors.STOP:
                self._publish(SharedVectors._CTRL + self.rank, seq)
                return
            self._execute(request)
            self._publish(SharedVectors._CTRL + self.rank, seq)

    def _request(self, request):
        """Rank 0: post, do the own share, wait for everybody."""
        ctrl = self.sv.ctrl
        self._seq += 1
        ctrl[1] = request
        self._publish(0, self._seq)     # published last, with release order
        if request != SharedVectors.STOP:
            self._execute(request)
        self._publish(SharedVectors._CTRL, self._seq)
        base = SharedVectors._CTRL
        self._wait(lambda: all(self._observe(base + r) == self._seq
                               for r in range(self.world)), self.TIMEOUT_S)

    def stop(self):
        if self.rank == 0 and self._seq >= 0:
            self._request(SharedVectors.STOP)
            self._seq = -1

    def close(self):
        self.stop()
        self.sv.close()

    # -- nlp.Evaluator surface (rank 0) ----------------------------------------
    def jac_structure(self):
        return self.problem.constr_jac_ind()

    def hess_structure(self):
        return self.problem.lag_hess_ind()

    def time_structure(self):
        from . import nlp
        return nlp.problem_time_structure(self.problem)

    def eval_fg(self, x):
        import time
        t0 = time.perf_counter()
        self.sv.dvec[:] = x
        self._request(self.X | 1 | 4)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        return float(self.sv.f[0]), self.sv.g.copy()

    def eval_all(self, x, sigma, lam):
        import time
        t0 = time.perf_counter()
        sv = self.sv
        sv.dvec[:] = x
        sv.lam[:] = lam
        sv.sigma = sigma
        self._request(self.X | self.LAMBDA | 31)
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        return (float(sv.f[0]), sv.grad.copy(), sv.g.copy(), sv.jac.copy(),
                sv.hess.copy())

    _GROUP = {1: 1 | 4, 4: 1 | 4, 2: 1 | 2 | 4 | 8, 8: 1 | 2 | 4 | 8, 16: 16}

    def ipopt_eval(self, which, x, new_x, out, sigma=None, lam=None):
        """One IPOPT callback (same grouping as ``nlp.GpuEvaluator``)."""
        import time
        t0 = time.perf_counter()
        sv = self.sv
        flags = 0
        if new_x or not self._have_x:
            sv.dvec[:] = x
            self._have_x, self._fresh, flags = True, 0, self.X
        if which == 16:
            sv.lam[:] = lam
            sv.sigma = sigma
            flags |= self.LAMBDA
            self._fresh &= ~16
        if not (self._fresh & which):
            mask = self._GROUP[which]
            self._request(flags | mask)
            self._fresh |= mask
            self.kernel_groups += 1
        name = dict(self.RESULTS)[which]
        out[...] = getattr(sv, name) if which != 1 else sv.f[0]
        self.seconds += time.perf_counter() - t0
        self.calls += 1


def solver_facing_evaluator(problem, rank, world, device=0, group=None,
                            reduce='peer'):
    """``SolverFacingEvaluator`` on the CUDA backend under an initialised
    ``torch.distributed`` process group (one process per GPU)."""
    import torch
    import torch.distributed as dist
    ev = ShardedEvaluator(problem, rank, world, device=device)
    # one explicit stream for the kernels, the copies and the NCCL calls
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    ev.handle.set_stream(stream.cuda_stream)
    hook = None
    if world > 1:
        mode = reduce
        if mode == 'peer' and not ev.agree_on_peer_reduce(group):
            mode = 'nccl'
        if mode != 'peer':
            ptr = ev.handle.device_ptrs()['reduce']

            class _Reduce:
                __cuda_array_interface__ = {
                    'shape': (ev.n_reduce,), 'typestr': '<f8',
                    'data': (int(ptr), False), 'version': 2}
            red = torch.as_tensor(_Reduce(), device=f'cuda:{device}')

            def hook(h):
                dist.all_reduce(red, group=group)
                h.apply_reduced(ptr)
    out = SolverFacingEvaluator(
        problem, ev.shard, ev.handle, rank, world,
        broadcast=lambda box: dist.broadcast_object_list(box, src=0,
                                                         group=group),
        barrier=lambda: dist.barrier(group), reduce_hook=hook,
        lib=ev.lib)
    out.sharded, out.stream = ev, stream
    return out
