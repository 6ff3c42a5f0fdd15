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
    def local_dvec(self, dvec):
        out = np.empty(self.loc.ndec)
        for i, v in enumerate(self.st.vars):
            c = v['core']
            lo, n = self.loc.var_off[i], self.loc.var_rows[i] * c
            if v['per_sample']:
                g0 = self.glob.var_off[i] + self.k0 * c
            else:
                g0 = self.glob.var_off[i]
            out[lo:lo + n] = dvec[g0:g0 + n]
        return out

    def local_multipliers(self, lam):
        out = np.empty(self.loc.ncons)
        for fi, f in enumerate(self.st.funs):
            if f['cons_index'] < 0:
                continue
            c = f['out_core']
            lo, n = self.loc.cons_off[fi], self.loc.fun_rows[fi] * c
            g0 = self.glob.cons_off[fi] + (self.k0 * c if f['per_sample']
                                           else 0)
            out[lo:lo + n] = lam[g0:g0 + n]
        return out

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

    def enable_peer_reduce(self, group=None):
        """Switch from ``all_reduce`` + ``cfem_apply_reduced`` to the fused
        in-kernel exchange over NVLink peer memory (``cfem_set_peers``).

        PyTorch is only the plumbing here: a symmetric-memory allocation
        (``torch.distributed._symmetric_memory``) gives every rank a device
        pointer to every other rank's inbox; the exchange itself is done by
        the last CTA of the per-sample kernel.
        """
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n_inbox, n_flag = self.handle.peer_layout(world)
        buf = symm.empty(n_inbox + n_flag, dtype=torch.float64,
                         device=torch.device('cuda', torch.cuda.current_device()))
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        torch.cuda.synchronize()
        dist.barrier(group)             # every inbox is zeroed before any store
        self.handle.set_peers(rank, world, ptrs,
                              [p + 8 * n_inbox for p in ptrs])
        self._peer = (buf, hdl)         # keep the mapping alive
        return True

    def set_point(self, dvec, obj_factor, lam):
        self.handle.set_dvec(self.shard.local_dvec(dvec))
        self.handle.set_multipliers(obj_factor,
                                    self.shard.local_multipliers(lam))
