"""NLP solver hand-off: ``problem.ipopt(dec_bounds, constr_bounds)``.

The reference scripts drive the solver through a context manager that yields
an object with ``add_str_option`` / ``add_num_option`` / ``add_int_option`` /
``set_scaling(obj_scale, dec_scale, constr_scale)`` / ``solve(dec0)``
(/root/reference/attas_sp_ml.py:153-159).  In the reference that object wraps
IPOPT's C interface (through the absent ``ceacoest`` / ``mseipopt`` packages).

Two drivers implement that surface here; both obtain every function value and
derivative from the CUDA path (:class:`GpuEvaluator`) -- there is no CPU
evaluation in this package:

* :class:`IpoptSolver` binds IPOPT's ``IpStdCInterface.h`` with ctypes when a
  ``libipopt`` can be found at run time (``CFEM_IPOPT_LIB`` or the loader
  path).  The KKT factorisation stays in IPOPT's own linear solver.
* :class:`InteriorPointSolver` is a self-contained primal-dual interior-point
  method (log barrier, fraction-to-boundary rule, l1-merit line search,
  curvature-based regularisation) that factorises the KKT matrix on the host
  with SciPy's sparse LU.  It exists because neither IPOPT nor HSL is present
  in the build image, and it is what the solution-parity tests and the
  Monte-Carlo driver run.  Every ``solve`` reports the time spent in callbacks
  (GPU path) and in the KKT factorisation (host) separately.
"""

import ctypes
import ctypes.util
import os
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import backend


# ----------------------------------------------------------------------------
# evaluators
# ----------------------------------------------------------------------------

class Evaluator:
    """What a solver needs from a problem (sizes, structure, callbacks)."""

    n = 0           # decision variables
    m = 0           # constraints

    def jac_structure(self):
        raise NotImplementedError

    def hess_structure(self):
        raise NotImplementedError

    def eval_fg(self, x):
        """(f, g) at x -- the line-search evaluation."""
        raise NotImplementedError

    def eval_all(self, x, sigma, lam):
        """(f, grad, g, jac values, hess values) at x."""
        raise NotImplementedError

    def time_structure(self):
        """(var_time [n], con_time [m]): sample index of every variable and
        constraint, -1 for parameters / parameter-only constraints; None if
        unknown.  Lets the KKT solver exploit the banded-in-time structure."""
        return None


def time_structure_of(decision, constraints, sample_vars, sample_cons):
    """Helper: time indices from ``{name: (offset, shape)}`` tables, the names
    of the per-sample variables and ``{name: shift}`` of the per-sample
    constraint functions.  ``shift`` is the largest row shift among the
    variables a function reads (1 for ``dynamics``, which reads ``xnext``):
    row k is then attributed to sample k + shift, the sample whose state it
    defines, so that every leading block of samples is a well-posed truncated
    problem (forward block elimination stays stable)."""
    n = max(off + int(np.prod(shape)) for off, shape in decision.values())
    m = max(off + int(np.prod(shape)) for off, shape in constraints.values())
    var_t = -np.ones(n, dtype=np.int64)
    con_t = -np.ones(m, dtype=np.int64)
    for table, names, out in ((decision, dict.fromkeys(sample_vars, 0), var_t),
                              (constraints, dict(sample_cons), con_t)):
        for name, shift in names.items():
            off, shape = table[name]
            rows = shape[0]
            core = int(np.prod(shape[1:]))
            out[off:off + rows * core] = np.repeat(np.arange(rows) + shift,
                                                   core)
    return var_t, con_t


def problem_time_structure(p):
    """``time_structure_of`` for an ``optim.Problem``."""
    st = p.structure
    dec = {n_: (s_.offset, s_.shape) for n_, s_ in p.decision.items()}
    con = {n_: (r.block.offset, r.block.shape)
           for n_, r in p.constraints.items()}
    svars = [v['name'] for v in st.vars if v['per_sample']]
    scons = {f['name']: max([r[2] for r in f['args'].values()
                             if r[0] == 'var'], default=0)
             for f in st.funs
             if f['per_sample'] and not f['is_objective']}
    return time_structure_of(dec, con, svars, scons)


class GpuEvaluator(Evaluator):
    """Callbacks served by the fused CUDA kernels of ``problem.backend``.

    ``eval_fg`` / ``eval_all`` (the built-in driver's calls) return VIEWS into
    one page-locked result block: they are valid until the next evaluation of
    this object (the driver consumes them at once).  ``ipopt_eval`` (IPOPT's
    five callbacks) moves the solver's own ``x`` / ``lambda`` / result arrays
    with no intermediate NumPy copy: the C ABI DMAs page-locked arrays directly
    and pipelines pageable ones through page-locked bounce buffers with a
    threaded host copy (``cfem_set_dvec`` / ``cfem_fetch``, csrc/cfem_host.inl).
    A solver whose buffers live as long as the problem can hand them to
    :meth:`pin` once; ``new_x`` (IPOPT passes it explicitly) avoids
    re-uploading an unchanged ``x``.
    """

    def __init__(self, problem):
        self.problem = problem
        self.be = problem.backend
        self.h = self.be.handle
        self.n, self.m = problem.ndec, problem.ncons
        self.buf = backend.HostBuffers(self.h)
        self._pinned = {}
        self.seconds = 0.0
        self.calls = 0
        self.kernel_groups = 0
        self._have_x = False
        self._fresh = 0

    def jac_structure(self):
        return self.problem.constr_jac_ind()

    def hess_structure(self):
        return self.problem.lag_hess_ind()

    def time_structure(self):
        return problem_time_structure(self.problem)

    # -- explicit page-locking of solver-owned arrays ---------------------------
    def pin(self, *arrays):
        """Page-lock caller-owned arrays (``cfem_host_register``) so that the
        DMA engines read / write them in place.  The caller guarantees that
        each allocation stays alive and in place until :meth:`unpin` /
        :meth:`close` -- never pin a buffer the solver may free or move."""
        lib = self.h.lib
        for arr in arrays:
            ptr, nbytes = arr.ctypes.data, arr.nbytes
            if nbytes == 0 or self._pinned.get(ptr, 0) >= nbytes:
                continue
            if ptr in self._pinned:
                lib.cfem_host_unregister(ptr)
                del self._pinned[ptr]
            if lib.cfem_host_register(ptr, nbytes) != 0:
                raise backend.CfemError('cfem_host_register failed for a '
                                        f'{nbytes}-byte array')
            self._pinned[ptr] = nbytes

    def unpin(self, *arrays):
        lib = self.h.lib
        ptrs = [a.ctypes.data for a in arrays] if arrays \
            else list(self._pinned)
        for ptr in ptrs:
            if ptr in self._pinned:
                lib.cfem_host_unregister(ptr)
                del self._pinned[ptr]

    def eval_fg(self, x):
        t0 = time.perf_counter()
        self.h.set_dvec(x)
        self._have_x, self._fresh = True, 0
        self.h.eval(backend.F | backend.G)
        self.h.fetch_async(backend.F, self.buf.f)
        self.h.fetch_async(backend.G, self.buf.g)
        self.h.synchronize()
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        return float(self.buf.f[0]), self.buf.g

    def eval_all(self, x, sigma, lam):
        t0 = time.perf_counter()
        self.h.set_dvec(x)
        self.h.set_multipliers(sigma, lam)
        self._have_x, self._fresh = True, 0
        self.h.eval(backend.ALL)
        self.buf.fetch_all()            # one D2H copy: [f | grad | g | jac | hess]
        self.seconds += time.perf_counter() - t0
        self.calls += 1
        b = self.buf
        return float(b.f[0]), b.grad, b.g, b.jac, b.hess

    # IPOPT-shaped single callbacks (used by IpoptSolver).  IPOPT asks for f
    # and g at every trial point and for grad f, the Jacobian and the Hessian
    # at accepted points, one callback at a time; the fused kernels serve a
    # whole group per launch and `new_x` tells when the groups go stale:
    #   f or g       -> one launch of mask F|G
    #   grad or Jac  -> one launch of mask F|GRAD|G|JAC
    #   Hessian      -> one launch of mask HESS (multipliers change per call)
    _GROUP = {backend.F: backend.F | backend.G,
              backend.G: backend.F | backend.G,
              backend.GRAD: backend.F | backend.GRAD | backend.G | backend.JAC,
              backend.JAC: backend.F | backend.GRAD | backend.G | backend.JAC,
              backend.HESS: backend.HESS}

    def ipopt_eval(self, which, x, new_x, out, sigma=None, lam=None):
        """One IPOPT callback: result ``which`` at ``x`` into the solver-owned
        array ``out`` (``sigma`` / ``lam`` for the Hessian)."""
        t0 = time.perf_counter()
        h = self.h
        if new_x or not self._have_x:
            h.set_dvec(x)               # from the solver's own x
            self._have_x = True
            self._fresh = 0
        if which == backend.HESS:
            h.set_multipliers(sigma, lam)
            self._fresh &= ~backend.HESS
        if not (self._fresh & which):
            mask = self._GROUP[which]
            h.eval(mask)
            self._fresh |= mask
            self.kernel_groups += 1
        h.fetch(which, out)             # into the solver's own array
        self.seconds += time.perf_counter() - t0
        self.calls += 1

    def close(self):
        self.unpin()
        self.buf.close()


# ----------------------------------------------------------------------------
# common solver surface
# ----------------------------------------------------------------------------

class Solver:
    """Option / scaling bookkeeping shared by both drivers."""

    def __init__(self, evaluator, dec_bounds, constr_bounds):
        self.ev = evaluator
        n, m = evaluator.n, evaluator.m
        dec_bounds = np.asarray(dec_bounds, dtype=float)
        constr_bounds = np.asarray(constr_bounds, dtype=float)
        if dec_bounds.shape != (2, n) or constr_bounds.shape != (2, m):
            raise ValueError(f'bounds must be (2, {n}) and (2, {m})')
        self.x_L, self.x_U = dec_bounds[0].copy(), dec_bounds[1].copy()
        self.g_L, self.g_U = constr_bounds[0].copy(), constr_bounds[1].copy()
        self.str_options, self.num_options, self.int_options = {}, {}, {}
        self.obj_scale = 1.0
        self.dec_scale = None
        self.constr_scale = None

    def add_str_option(self, key, value):
        self.str_options[key] = value

    def add_num_option(self, key, value):
        self.num_options[key] = float(value)

    def add_int_option(self, key, value):
        self.int_options[key] = int(value)

    def set_scaling(self, obj_scale, dec_scale, constr_scale):
        """IPOPT's user scaling (attas_sp_ml.py:158): the solver works on
        ``obj_scale*f``, ``dec_scale*x`` and ``constr_scale*g``."""
        self.obj_scale = float(obj_scale)
        self.dec_scale = None if dec_scale is None else \
            np.array(dec_scale, dtype=float)
        self.constr_scale = None if constr_scale is None else \
            np.array(constr_scale, dtype=float)

    def solve(self, dec0):
        raise NotImplementedError

    def close(self):
        close = getattr(self.ev, 'close', None)
        if close:
            close()


# ----------------------------------------------------------------------------
# built-in interior-point driver
# ----------------------------------------------------------------------------

def _bt_negative_py(D, E):
    """Negative-eigenvalue count of a symmetric block-tridiagonal matrix
    (diagonal blocks D[i], sub-diagonal blocks E[i]) by the Schur recursion
    S_{i+1} = D_{i+1} - E_i S_i^-1 E_i'; -1 if a pivot block is numerically
    singular."""
    neg = 0
    S = D[0].copy()
    nblk = D.shape[0]
    for i in range(nblk):
        S = 0.5 * (S + S.T)
        ev_ = np.linalg.eigvalsh(S)
        big = max(1.0, np.max(np.abs(ev_)))
        if np.min(np.abs(ev_)) <= 1e-13 * big:
            return -1
        neg += int((ev_ < 0).sum())
        if i + 1 < nblk:
            S = D[i + 1] - E[i] @ np.linalg.solve(S, E[i].T)
    return neg


try:        # the recursion is sequential over samples: compile it if we can
    import numba as _numba
    _bt_negative = _numba.njit(cache=True)(_bt_negative_py)
except Exception:       # pragma: no cover - numba is optional
    _bt_negative = _bt_negative_py


def _sym_factor(S):
    """Bunch-Kaufman ``S = L D L'`` (LAPACK dsytrf) of a dense symmetric
    matrix: ``(ldu, piv, neg)`` for ``dsytrs`` plus the number of negative
    eigenvalues (Sylvester: the inertia of the 1x1 / 2x2 blocks of D), ``neg``
    None when D is numerically singular; None if the factorisation broke
    down.  One O(n^3/3) pass instead of an LU for the solve and a full
    eigenvalue decomposition for the inertia."""
    if S.shape[0] == 0:
        return np.zeros((0, 0)), np.zeros(0, dtype=np.int32), 0
    lwork, _ = sla.lapack.dsytrf_lwork(S.shape[0], lower=1)     # blocked code
    ldu, piv, info = sla.lapack.dsytrf(S, lower=1, lwork=max(1, int(lwork)))
    if info != 0:
        return None
    d, sub = np.diag(ldu), np.diag(ldu, -1)
    two = piv < 0                       # both rows of a 2x2 block are flagged
    first = np.zeros(len(piv), dtype=bool)
    k, n = 0, len(piv)
    while k < n:                        # leading rows of the 2x2 blocks
        if two[k]:
            first[k] = True
            k += 2
        else:
            k += 1
    ev1 = d[~two]
    lead = np.flatnonzero(first)
    a, c, b = d[lead], d[lead + 1], sub[lead]
    half, disc = 0.5 * (a + c), np.sqrt(0.25 * (a - c) ** 2 + b * b)
    ev = np.concatenate([ev1, half - disc, half + disc])
    if not np.all(np.isfinite(ev)) or \
            np.min(np.abs(ev)) <= 1e-14 * max(1.0, np.max(np.abs(ev))):
        return ldu, piv, None
    return ldu, piv, int((ev < 0).sum())


class BorderedBandKKT:
    """Host solver for the KKT systems of the filter-error problems.

    The KKT matrix [[H + dw I, J'], [J, -dc I]] of a collocation problem is
    block-banded along time (each sample couples to its neighbours only,
    /root/reference/fem.py:47-52) plus a thin dense border: the model
    parameters, which every sample's constraints touch.  A general sparse LU
    loses that structure as soon as it pivots (measured: 4.4 s at N = 2000),
    so the border is split off explicitly:

        [Kbb Kbp] [zb]   [rb]      Kbb: banded (reverse Cuthill-McKee order),
        [Kpb Kpp] [zp] = [rp]           LU with partial pivoting inside the band
                                   S = Kpp - Kpb Kbb^-1 Kbp: small, dense

    Cost is linear in the number of samples.  The border is found from the
    sparsity pattern alone (vertices whose degree is far above the median).
    SciPy offers no sparse LDL^T; the inertia that IPOPT's regularisation
    heuristic needs is obtained from the same splitting (Haynsworth):
    inertia(K) = inertia(Kbb) + inertia(S), with inertia(Kbb) from a
    block-tridiagonal LDL^T of the banded block (``_band_negative``).
    """

    def __init__(self, var_time=None, con_time=None):
        """``var_time`` / ``con_time``: sample index of every (free) variable
        and constraint, -1 for parameters and parameter-only constraints.
        Without them the border is guessed from vertex degrees and no inertia
        is available."""
        self._key = None
        self.var_time, self.con_time = var_time, con_time
        self.blocks = None

    def _analyse(self, K):
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        if self.var_time is not None:
            t = np.concatenate([self.var_time, self.con_time])
            self.border = np.flatnonzero(t < 0)
            inner = np.flatnonzero(t >= 0)
            order = np.argsort(t[inner], kind='stable')
            self.inner = inner[order]
            tt = t[self.inner]
            # block boundaries: one block per sample
            starts = np.flatnonzero(np.r_[True, tt[1:] != tt[:-1]])
            self.blocks = np.r_[starts, len(tt)]
            return
        A = (abs(K) + abs(K).T).tocsr()
        deg = np.diff(A.indptr)
        cut = max(64, 8 * int(np.median(deg)))
        border = np.flatnonzero(deg > cut)
        inner = np.flatnonzero(deg <= cut)
        if len(inner):
            sub = A[inner][:, inner].tocsr()
            inner = inner[reverse_cuthill_mckee(sub, symmetric_mode=True)]
        self.inner, self.border = inner, border

    def _band_negative(self, Kbb):
        """Number of negative eigenvalues of the time-banded block by
        block-tridiagonal LDL^T.  One block per sample holds that sample's
        states, innovations and the multipliers of its defect / innovation
        rows -- itself a small KKT matrix with a square constraint Jacobian,
        hence nonsingular -- and a sample couples only to its neighbours
        (/root/reference/fem.py:47-52), so the Schur recursion
        S_{k+1} = D_{k+1} - E_k S_k^-1 E_k' is a congruence and, by Sylvester,
        the inertia is the sum of the inertias of the S_k.  Returns None when
        no time structure is known or a pivot block is numerically singular."""
        if self.blocks is None:
            return None
        nb_ = Kbb.shape[0]
        if nb_ == 0:
            return 0
        bounds = self.blocks
        nblk = len(bounds) - 1
        sizes = np.diff(bounds)
        w = int(sizes.max())
        blk_of = np.repeat(np.arange(nblk), sizes)
        pos_in = np.arange(nb_) - bounds[blk_of]
        coo = Kbb.tocoo()
        bi, bj = blk_of[coo.row], blk_of[coo.col]
        if np.any(np.abs(bi - bj) > 1):
            return None             # not block tridiagonal in time
        ri, rj = pos_in[coo.row], pos_in[coo.col]
        D = np.zeros((nblk, w, w))
        E = np.zeros((max(nblk - 1, 1), w, w))
        dmask = bi == bj            # (a CSR-derived COO has no duplicates)
        D[bi[dmask], ri[dmask], rj[dmask]] = coo.data[dmask]
        emask = bi == bj + 1
        E[bj[emask], ri[emask], rj[emask]] = coo.data[emask]
        for i in np.flatnonzero(sizes < w):     # identity padding
            idx = np.arange(sizes[i], w)
            D[i, idx, idx] = 1.0
        neg = _bt_negative(D, np.ascontiguousarray(E))
        return None if neg < 0 else int(neg)

    def solve(self, H, J, delta_w, delta_c, rhs, want_inertia=False):
        n, m = H.shape[0], J.shape[0]
        blocks = [[H + delta_w * sp.identity(n), J.T],
                  [J, -delta_c * sp.identity(m) if delta_c > 0 else None]]
        K = sp.bmat(blocks, format='csr')
        key = (n, m, H.nnz, J.nnz)
        if key != self._key:
            self._analyse(K)
            self._key = key
        inner, border = self.inner, self.border
        neg = None
        self._factors = None
        try:
            Kb = K[inner]
            Kbb = Kb[:, inner]
            lu = spla.splu(Kbb.tocsc(), permc_spec='NATURAL')
            Y = cpl = Kpb = None
            if len(border):
                # only the border columns that touch the samples (the model
                # parameters proper) need a banded solve
                Kbp_s = Kb[:, border].tocsc()
                cpl = np.flatnonzero(np.diff(Kbp_s.indptr))
                Kp = K[border]
                Kpb = Kp[:, inner]
                S = Kp[:, border].toarray()
                if len(cpl):
                    Y = lu.solve(Kbp_s[:, cpl].toarray())
                    S[:, cpl] -= Kpb @ Y
                # the border complement is symmetric up to rounding: ONE
                # Bunch-Kaufman LDL^T serves the solve and the inertia
                S_lu = _sym_factor(0.5 * (S + S.T))
                if S_lu is None:
                    return None         # singular border complement
            else:
                S = np.zeros((0, 0))
                S_lu = None
            # kept for further right-hand sides with the same matrix
            # (iterative refinement): no second factorisation
            self._factors = (K, lu, Y, cpl, Kpb, S_lu)
            zb, zp = self._back_solve(rhs)
            if not (np.all(np.isfinite(zb)) and np.all(np.isfinite(zp))):
                self._factors = None
                return None             # singular border complement
            if want_inertia:
                neg = self._band_negative(Kbb)
                if neg is not None and len(border):
                    neg = None if S_lu[2] is None else neg + S_lu[2]
        except (RuntimeError, np.linalg.LinAlgError):
            return None
        sol = np.empty(n + m)
        sol[inner], sol[border] = zb, zp
        return sol, K, neg

    def _back_solve(self, rhs):
        _, lu, Y, cpl, Kpb, S_lu = self._factors
        inner, border = self.inner, self.border
        zb = lu.solve(rhs[inner])
        if S_lu is None:
            return zb, np.zeros(0)
        zp, info = sla.lapack.dsytrs(S_lu[0], S_lu[1], rhs[border] - Kpb @ zb,
                                     lower=1)
        if info != 0:
            raise np.linalg.LinAlgError('dsytrs failed')
        if Y is not None:
            zb = zb - Y @ zp[cpl]
        return zb, zp

    def resolve(self, rhs):
        """Solution for another right-hand side with the matrix of the latest
        ``solve`` (its factors are reused), or None."""
        if getattr(self, '_factors', None) is None:
            return None
        try:
            zb, zp = self._back_solve(rhs)
        except (RuntimeError, np.linalg.LinAlgError, ValueError):
            return None
        sol = np.empty(len(rhs))
        sol[self.inner], sol[self.border] = zb, zp
        return sol


def _kkt_solve(solver, H, J, delta_w, delta_c, rhs, want_inertia=False):
    """(sol, inertia_ok) with one step of iterative refinement.
    ``inertia_ok``: True when the matrix has exactly m negative eigenvalues,
    False when not (or singular), None when inertia was not requested."""
    out = solver.solve(H, J, delta_w, delta_c, rhs, want_inertia)
    if out is None:
        return None, False
    sol, K, neg = out
    res = rhs - K @ sol
    if np.all(np.isfinite(res)) and np.linalg.norm(res) > \
            1e-13 * np.linalg.norm(rhs):
        if hasattr(solver, 'resolve'):
            cor = solver.resolve(res)           # same factors, new rhs
        else:
            cor = solver.solve(H, J, delta_w, delta_c, res)
            cor = None if cor is None else cor[0]
        if cor is not None and np.all(np.isfinite(cor)):
            sol = sol + cor
    if not want_inertia or neg is None:
        return sol, None
    return sol, neg == J.shape[0]


class InteriorPointSolver(Solver):
    """Primal-dual interior point method for

        min  s_f f(x)   s.t.  g(x) = g_L (= g_U),   x_L <= x <= x_U

    (all constraints of the reference's problems are equalities; variables with
    ``x_L == x_U`` are eliminated like IPOPT's ``make_parameter``).  The
    iteration follows the IPOPT paper (Waechter & Biegler 2006): monotone
    barrier update, fraction-to-boundary rule, inertia-correcting
    regularisation, filter line search with second-order correction, scaled
    optimality error; the restoration phase is a plain minimum-norm
    feasibility iteration.
    """

    name = 'builtin-ipm'

    # IPOPT's default constants
    GAMMA_THETA, GAMMA_PHI, ETA_PHI = 1e-5, 1e-8, 1e-8
    S_THETA, S_PHI, DELTA_SW = 1.1, 2.3, 1.0
    KAPPA_SOC, MAX_SOC = 0.99, 4

    def solve(self, dec0):
        """Run to completion with this solver's own evaluator."""
        steps = self.solve_steps(dec0)
        request = next(steps)
        while True:
            try:
                if request[0] == 'all':
                    result = self.ev.eval_all(*request[1:])
                else:
                    result = self.ev.eval_fg(request[1])
                request = steps.send(result)
            except StopIteration as stop:
                return stop.value

    def solve_steps(self, dec0):
        """The iteration as a generator: yields evaluation requests
        ``('all', x, sigma, lam)`` / ``('fg', x)`` and is sent the results
        (``Evaluator.eval_all`` / ``eval_fg`` tuples); returns
        ``(x_opt, info)``.  A batch driver can advance many solvers in
        lock-step and serve their requests with ONE batched kernel launch
        (``fit.BatchFitter``)."""
        ev = self.ev
        n, m = ev.n, ev.m
        tol = self.num_options.get('tol', 1e-8)
        max_iter = self.int_options.get('max_iter', 3000)
        verbose = self.int_options.get('print_level', 0) > 0
        if np.any(self.g_L != self.g_U):
            raise NotImplementedError('inequality constraints are not used by '
                                      'the reference problems')
        sf = self.obj_scale
        dx_s = np.ones(n) if self.dec_scale is None else self.dec_scale
        dc_s = np.ones(m) if self.constr_scale is None else self.constr_scale

        x_full = np.array(dec0, dtype=float)
        fixed = self.x_L == self.x_U
        x_full[fixed] = self.x_L[fixed]
        free = np.flatnonzero(~fixed)
        nf = len(free)
        pos = -np.ones(n, dtype=np.int64)
        pos[free] = np.arange(nf)
        s = dx_s[free]
        with np.errstate(invalid='ignore'):
            lo, hi = self.x_L[free] * s, self.x_U[free] * s
        neg = s < 0
        lo[neg], hi[neg] = hi[neg].copy(), lo[neg].copy()
        has_lo, has_hi = np.isfinite(lo), np.isfinite(hi)
        any_bounds = bool(has_lo.any() or has_hi.any())
        lo_f = np.where(has_lo, lo, 0.0)
        hi_f = np.where(has_hi, hi, 0.0)

        jr, jc = ev.jac_structure()
        hr, hc = ev.hess_structure()
        jkeep = ~fixed[jc]
        hkeep = ~(fixed[hr] | fixed[hc])
        jr_f, jc_f = jr[jkeep], pos[jc[jkeep]]
        hr_f, hc_f = pos[hr[hkeep]], pos[hc[hkeep]]
        jscale = dc_s[jr[jkeep]] / dx_s[jc[jkeep]]
        hscale = 1.0 / (dx_s[hr[hkeep]] * dx_s[hc[hkeep]])
        offdiag = hr_f != hc_f

        ts = ev.time_structure()
        self._kkt_solver = BorderedBandKKT() if ts is None else \
            BorderedBandKKT(ts[0][free], ts[1])
        self._t_lin = 0.0
        t_start = time.perf_counter()
        ev_t0, ev_c0 = getattr(ev, 'seconds', 0.0), getattr(ev, 'calls', 0)

        def unscale(xs):
            x_full[free] = xs / s
            return x_full

        def evaluate(xs, lam_s):
            f, grad, g, jv, hv = yield ('all', unscale(xs).copy(), sf,
                                        dc_s * lam_s)
            J = sp.csr_matrix((jv[jkeep] * jscale, (jr_f, jc_f)),
                              shape=(m, nf))
            hvals = hv[hkeep] * hscale
            W = sp.coo_matrix((hvals, (hr_f, hc_f)), shape=(nf, nf)) \
                + sp.coo_matrix((hvals[offdiag],
                                 (hc_f[offdiag], hr_f[offdiag])),
                                shape=(nf, nf))
            return (sf * f, sf * grad[free] / s, dc_s * (g - self.g_L), J,
                    W.tocsr())

        def evaluate_fc(xs):
            f, g = yield ('fg', unscale(xs).copy())
            return sf * f, dc_s * (g - self.g_L)

        def slacks(x):
            return (np.where(has_lo, x - lo_f, 1.0),
                    np.where(has_hi, hi_f - x, 1.0))

        def barrier(x, f, mu_):
            if not any_bounds:
                return f
            sl, su = slacks(x)
            if np.any(sl <= 0) or np.any(su <= 0):
                return np.inf
            return f - mu_ * (np.log(sl[has_lo]).sum()
                              + np.log(su[has_hi]).sum())

        def opt_error(mu_, gs, cs, J, lam, zl, zu, x):
            sl, su = slacks(x)
            dual = gs + J.T @ lam - zl + zu
            s_max = 100.0
            nb = int(has_lo.sum() + has_hi.sum())
            s_d = max(s_max, (np.abs(lam).sum() + zl.sum() + zu.sum())
                      / max(1, m + nb)) / s_max
            s_c = max(s_max, (zl.sum() + zu.sum()) / max(1, nb)) / s_max
            comp = 0.0
            if any_bounds:
                comp = max(np.max(np.abs(sl * zl - mu_)[has_lo], initial=0.0),
                           np.max(np.abs(su * zu - mu_)[has_hi], initial=0.0))
            return max(np.max(np.abs(dual), initial=0.0) / s_d,
                       np.max(np.abs(cs), initial=0.0), comp / s_c)

        def max_step(v, dv, mask, tau):
            idx = mask & (dv < 0)
            if not idx.any():
                return 1.0
            return min(1.0, float(np.min(-tau * v[idx] / dv[idx])))

        # ---- initial point (IPOPT bound_push / bound_frac)
        xs = x_full[free] * s
        if any_bounds:
            k1 = k2 = self.num_options.get('bound_push', 1e-2)
            span = np.where(has_lo & has_hi, hi_f - lo_f, np.inf)
            pl = np.minimum(k1 * np.maximum(1, np.abs(lo_f)), k2 * span)
            pu = np.minimum(k1 * np.maximum(1, np.abs(hi_f)), k2 * span)
            xs = np.where(has_lo, np.maximum(xs, lo_f + pl), xs)
            xs = np.where(has_hi, np.minimum(xs, hi_f - pu), xs)
        zl = np.where(has_lo, 1.0, 0.0)
        zu = np.where(has_hi, 1.0, 0.0)
        lam = np.zeros(m)
        mu = self.num_options.get('mu_init', 0.1) if any_bounds else 0.0

        fs, gs, cs, J, W = yield from evaluate(xs, lam)
        sol, _ = self._kkt(sp.identity(nf, format='csr'), J, 0.0, 0.0,
                           np.concatenate([-(gs - zl + zu), np.zeros(m)]))
        if sol is not None and np.all(np.isfinite(sol)) \
                and np.max(np.abs(sol[nf:]), initial=0.0) <= 1e3:
            lam = sol[nf:]
            fs, gs, cs, J, W = yield from evaluate(xs, lam)

        theta0 = np.abs(cs).sum()
        theta_max = self.num_options.get('theta_max_fact', 1e2) \
            * max(1.0, theta0)
        theta_min = 1e-4 * max(1.0, theta0)
        filt = []                     # list of (theta, phi) corners

        def in_filter(th, ph):
            if th >= theta_max:
                return True
            return any(th >= t and ph >= p_ for t, p_ in filt)

        delta_last = 0.0
        status = 'max_iter'
        it = 0
        err0 = np.inf
        ls_note = ''
        for it in range(max_iter + 1):
            err0 = opt_error(0.0, gs, cs, J, lam, zl, zu, xs)
            if verbose:
                print(f'{it:4d} f={fs:+.10e} theta={np.abs(cs).sum():.2e} '
                      f'err={err0:.2e} mu={mu:.1e} reg={delta_last:.1e} '
                      f'{ls_note}')
            if err0 <= tol:
                status = 'solved'
                break
            if it == max_iter:
                break
            # barrier parameter update (IPOPT eq. 7) -- resets the filter
            while any_bounds and mu > tol / 10 and \
                    opt_error(mu, gs, cs, J, lam, zl, zu, xs) <= 10 * mu:
                mu = max(tol / 10, min(0.2 * mu, mu ** 1.5))
                filt = []
            tau = max(0.99, 1 - mu)

            sl, su = slacks(xs)
            sig = np.where(has_lo, zl / sl, 0.0) \
                + np.where(has_hi, zu / su, 0.0)
            gphi = gs - np.where(has_lo, mu / sl, 0.0) \
                + np.where(has_hi, mu / su, 0.0)
            Hd = (W + sp.diags(sig)).tocsr()
            rhs = -np.concatenate([gphi + J.T @ lam, cs])

            # ---- search direction with inertia correction (algorithm IC)
            delta = 0.0
            step = None
            for attempt in range(60):
                sol, inertia_ok = self._kkt(Hd, J, delta,
                                            1e-8 if delta > 0 else 0.0, rhs,
                                            want_inertia=True)
                ok = sol is not None and np.all(np.isfinite(sol))
                if ok and inertia_ok is None:       # inertia-free test
                    d = sol[:nf]
                    curv = d @ (Hd @ d) + delta * (d @ d)
                    ok = curv >= 1e-8 * (d @ d)
                elif ok:
                    ok = inertia_ok
                if ok:
                    step = sol
                    break
                if delta == 0.0:
                    delta = 1e-4 if delta_last == 0 else max(1e-20,
                                                             delta_last / 3)
                else:
                    delta *= 8.0 if delta_last > 0 else 100.0
                if delta > 1e40:
                    break
            if step is None:
                status = 'kkt_failure'
                break
            if delta > 0:
                delta_last = delta
            d, dlam = step[:nf], step[nf:]

            # ---- filter line search (IPOPT algorithm A, steps A-5.x)
            theta = np.abs(cs).sum()
            phi = barrier(xs, fs, mu)
            dphi = gphi @ d
            a_max = min(max_step(sl, d, has_lo, tau),
                        max_step(su, -d, has_hi, tau))
            switching = dphi < 0
            if switching and theta > 0:
                a_min = min(self.GAMMA_THETA,
                            self.GAMMA_PHI * theta / (-dphi),
                            self.DELTA_SW * theta ** self.S_THETA
                            / (-dphi) ** self.S_PHI)
            elif switching:
                a_min = self.GAMMA_THETA
            else:
                a_min = self.GAMMA_THETA
            a_min *= 0.05

            def acceptable(alpha, xt, ft, ct):
                th_t = np.abs(ct).sum()
                ph_t = barrier(xt, ft, mu)
                if not np.isfinite(ph_t) or in_filter(th_t, ph_t):
                    return False, False
                sw = (switching and theta <= theta_min and
                      alpha * (-dphi) ** self.S_PHI
                      > self.DELTA_SW * theta ** self.S_THETA)
                if sw:
                    return (ph_t <= phi + self.ETA_PHI * alpha * dphi
                            + 10 * np.finfo(float).eps * abs(phi)), True
                ok_ = (th_t <= (1 - self.GAMMA_THETA) * theta or
                       ph_t <= phi - self.GAMMA_PHI * theta)
                return ok_, False

            alpha = a_max
            accepted = None
            first = True
            ls_note = ''
            while alpha >= a_min:
                xt = xs + alpha * d
                ft, ct = yield from evaluate_fc(xt)
                ok, armijo = acceptable(alpha, xt, ft, ct)
                if ok:
                    accepted = (alpha, xt, d, dlam, armijo)
                    break
                # second-order correction on the first trial point
                if first and np.abs(ct).sum() >= theta:
                    c_soc = alpha * cs + ct
                    th_old = theta
                    for p_ in range(self.MAX_SOC):
                        rhs_soc = -np.concatenate([gphi + J.T @ lam, c_soc])
                        sol, _ = self._kkt(Hd, J, delta,
                                           1e-8 if delta > 0 else 0.0, rhs_soc)
                        if sol is None or not np.all(np.isfinite(sol)):
                            break
                        dc_, dlc = sol[:nf], sol[nf:]
                        a_soc = min(max_step(sl, dc_, has_lo, tau),
                                    max_step(su, -dc_, has_hi, tau))
                        xt = xs + a_soc * dc_
                        ft, ct = yield from evaluate_fc(xt)
                        ok, armijo = acceptable(a_soc, xt, ft, ct)
                        if ok:
                            accepted = (a_soc, xt, dc_, dlc, armijo)
                            ls_note = f'soc{p_ + 1}'
                            break
                        th_soc = np.abs(ct).sum()
                        if th_soc > self.KAPPA_SOC * th_old:
                            break
                        th_old = th_soc
                        c_soc = a_soc * c_soc + ct
                    if accepted:
                        break
                first = False
                alpha *= 0.5
            if accepted is None:
                # ---- feasibility restoration (minimum-norm constraint steps)
                filt.append(((1 - self.GAMMA_THETA) * theta,
                             phi - self.GAMMA_PHI * theta))
                # Levenberg-Marquardt on 0.5*|c|^2 within the bounds:
                #   [lm*I  J'] [d]   [ 0]
                #   [J    -I ] [r] = [-c]   <=>  d = -(J'J + lm I)^-1 J'c
                restored = False
                xr, cr, Jr = xs.copy(), cs, J
                fs_r, gs_r, W_r = fs, gs, W
                zeros_h = sp.csr_matrix((nf, nf))
                lm = 1e-2
                r_it = 0
                for r_it in range(200):
                    sol, _ = self._kkt(zeros_h, Jr, lm, 1.0,
                                       np.concatenate([np.zeros(nf), -cr]))
                    if sol is None or not np.all(np.isfinite(sol)):
                        lm *= 10.0
                        if lm > 1e12:
                            break
                        continue
                    dr = sol[:nf]
                    slr, sur = slacks(xr)
                    ar = min(max_step(slr, dr, has_lo, tau),
                             max_step(sur, -dr, has_hi, tau))
                    xt = xr + ar * dr
                    ft, ct = yield from evaluate_fc(xt)
                    if ct @ ct < (1 - 1e-4 * ar) * (cr @ cr):
                        xr = xt
                        lm = max(lm / 3.0, 1e-10)
                        fs_r, gs_r, cr, Jr, W_r = yield from evaluate(xr, lam)
                        th_r = np.abs(cr).sum()
                        if th_r <= 0.9 * theta and \
                                not in_filter(th_r, barrier(xr, fs_r, mu)):
                            restored = True
                            break
                    else:
                        lm *= 4.0
                        if lm > 1e12:
                            break
                if not restored:
                    status = ('solved_to_acceptable_level'
                              if err0 <= 1e3 * tol else 'restoration_failure')
                    break
                xs = xr
                ls_note = f'restoration({r_it + 1})'
                fs, gs, cs, J, W = fs_r, gs_r, cr, Jr, W_r
                # multipliers: least squares at the restored point
                sol, _ = self._kkt(sp.identity(nf, format='csr'), J, 0.0, 0.0,
                                   np.concatenate([-(gs - zl + zu),
                                                   np.zeros(m)]))
                if sol is not None and np.all(np.isfinite(sol)) and \
                        np.max(np.abs(sol[nf:]), initial=0.0) <= 1e3:
                    lam = sol[nf:]
                else:
                    lam = np.zeros(m)
                fs, gs, cs, J, W = yield from evaluate(xs, lam)
                continue

            alpha, xt, d_acc, dlam_acc, armijo = accepted
            if not armijo:
                filt.append(((1 - self.GAMMA_THETA) * theta,
                             phi - self.GAMMA_PHI * theta))
            ls_note = (ls_note + f' a={alpha:.2e}').strip()
            dzl = np.where(has_lo, mu / sl - zl - zl / sl * d_acc, 0.0)
            dzu = np.where(has_hi, mu / su - zu + zu / su * d_acc, 0.0)
            a_d = min(max_step(zl, dzl, has_lo, tau),
                      max_step(zu, dzu, has_hi, tau))
            xs = xt
            lam = lam + alpha * dlam_acc
            zl = zl + a_d * dzl
            zu = zu + a_d * dzu
            if any_bounds:          # keep z near the central path (eq. 16)
                sl, su = slacks(xs)
                zl = np.where(has_lo, np.clip(zl, mu / (1e10 * sl),
                                              1e10 * mu / sl), 0.0)
                zu = np.where(has_hi, np.clip(zu, mu / (1e10 * su),
                                              1e10 * mu / su), 0.0)
            fs, gs, cs, J, W = yield from evaluate(xs, lam)

        total = time.perf_counter() - t_start
        x_opt = unscale(xs).copy()
        mult_x_L = np.zeros(n)
        mult_x_U = np.zeros(n)
        mult_x_L[free] = zl * np.abs(s) / abs(sf)
        mult_x_U[free] = zu * np.abs(s) / abs(sf)
        info = {
            'status': status, 'solver': self.name, 'iterations': it,
            'obj': fs / sf, 'scaled_obj': fs, 'error': err0,
            'mult_g': dc_s * lam / sf,
            'mult_x_L': mult_x_L, 'mult_x_U': mult_x_U,
            'seconds_total': total, 'seconds_kkt': self._t_lin,
            'seconds_callbacks': getattr(ev, 'seconds', 0.0) - ev_t0,
            'callback_calls': getattr(ev, 'calls', 0) - ev_c0,
        }
        return x_opt, info

    def _kkt(self, H, J, delta_w, delta_c, rhs, want_inertia=False):
        t0 = time.perf_counter()
        out = _kkt_solve(self._kkt_solver, H, J, delta_w, delta_c, rhs,
                         want_inertia)
        self._t_lin += time.perf_counter() - t0
        return out


# ----------------------------------------------------------------------------
# IPOPT through its C interface
# ----------------------------------------------------------------------------

def find_ipopt():
    """Path of a loadable libipopt, or None."""
    cand = [os.environ.get('CFEM_IPOPT_LIB'), ctypes.util.find_library('ipopt'),
            'libipopt.so', 'libipopt.so.3', 'libipopt.so.1']
    for c in cand:
        if not c:
            continue
        try:
            ctypes.CDLL(c)
            return c
        except OSError:
            continue
    return None


_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
# Bool (*)(Index n, Number* x, Bool new_x, ...) -- IpStdCInterface.h
_EVAL_F = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp,
                           ctypes.c_void_p)
_EVAL_GRAD_F = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                                _dp, ctypes.c_void_p)
_EVAL_G = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                           ctypes.c_int, _dp, ctypes.c_void_p)
_EVAL_JAC_G = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int, _ip, _ip, _dp,
                               ctypes.c_void_p)
_EVAL_H = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                           ctypes.c_double, ctypes.c_int, _dp, ctypes.c_int,
                           ctypes.c_int, _ip, _ip, _dp, ctypes.c_void_p)


class IpoptSolver(Solver):
    """IPOPT (``CreateIpoptProblem`` / ``IpoptSolve``) driven by the CUDA path.

    ``x``, ``g``, ``grad_f`` and ``values`` are IPOPT-owned host buffers, valid
    only during a callback; results are copied straight into them
    (``cfem_fetch``: pageable arrays are pipelined through page-locked bounce
    buffers by a threaded copy, no assumption on the pointers).  IPOPT's
    ``TNLPAdapter`` allocates ``x``, ``g``, ``grad_f`` and the Jacobian values
    once per problem; ``pin_buffers`` (``CFEM_PIN_IPOPT_BUFFERS=1``, off by
    default) page-locks those on first sight so that they are DMA targets
    themselves -- only for IPOPT builds known to keep them in place; the
    Hessian values array (a fresh matrix per iterate) is never pinned.  A CUDA
    failure is reported to IPOPT as an evaluation error (callback returns
    FALSE).
    """

    name = 'ipopt'

    def __init__(self, evaluator, dec_bounds, constr_bounds, libpath=None):
        super().__init__(evaluator, dec_bounds, constr_bounds)
        libpath = libpath or find_ipopt()
        if not libpath:
            raise RuntimeError('no libipopt found (set CFEM_IPOPT_LIB)')
        self.lib = lib = ctypes.CDLL(libpath)
        lib.CreateIpoptProblem.restype = ctypes.c_void_p
        lib.CreateIpoptProblem.argtypes = [
            ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, _EVAL_F, _EVAL_G, _EVAL_GRAD_F,
            _EVAL_JAC_G, _EVAL_H]
        lib.FreeIpoptProblem.argtypes = [ctypes.c_void_p]
        lib.AddIpoptStrOption.argtypes = [ctypes.c_void_p, ctypes.c_char_p,
                                          ctypes.c_char_p]
        lib.AddIpoptNumOption.argtypes = [ctypes.c_void_p, ctypes.c_char_p,
                                          ctypes.c_double]
        lib.AddIpoptIntOption.argtypes = [ctypes.c_void_p, ctypes.c_char_p,
                                          ctypes.c_int]
        lib.SetIpoptProblemScaling.argtypes = [ctypes.c_void_p,
                                               ctypes.c_double, _dp, _dp]
        lib.IpoptSolve.restype = ctypes.c_int
        lib.IpoptSolve.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, _dp,
                                   _dp, ctypes.c_void_p]
        ev = evaluator
        n, m = ev.n, ev.m
        jr, jc = ev.jac_structure()
        hr, hc = ev.hess_structure()
        if max(len(jr), len(hr), n, m) >= 2 ** 31:
            raise OverflowError(
                "IPOPT's Index is a 32-bit int: n, m, nnz_jac and nnz_hess "
                f'must stay below 2**31 (got n={n}, m={m}, nnz_jac={len(jr)}, '
                f'nnz_hess={len(hr)}); shorten the trajectory or build IPOPT '
                'with 64-bit indices')
        self.pin_buffers = os.environ.get('CFEM_PIN_IPOPT_BUFFERS') == '1'
        pin = getattr(ev, 'pin', None)

        def stable(*arrays):
            if self.pin_buffers and pin is not None:
                pin(*arrays)
            return arrays
        self._jr, self._jc = jr.astype(np.int32), jc.astype(np.int32)
        self._hr, self._hc = hr.astype(np.int32), hc.astype(np.int32)

        def arr(ptr, count):
            return np.ctypeslib.as_array(ptr, shape=(count,))

        def guard(fn):
            def wrapped(*args):
                try:
                    fn(*args)
                    return 1
                except Exception as exc:        # -> IPOPT "evaluation error"
                    self.last_error = exc
                    return 0
            return wrapped

        @guard
        def eval_f(n_, x, new_x, obj, _):
            out = np.empty(1)
            ev.ipopt_eval(backend.F, arr(x, n), new_x, out)
            obj[0] = out[0]

        @guard
        def eval_grad_f(n_, x, new_x, grad, _):
            ev.ipopt_eval(backend.GRAD, *stable(arr(x, n)), new_x,
                          *stable(arr(grad, n)))

        @guard
        def eval_g(n_, x, new_x, m_, g, _):
            ev.ipopt_eval(backend.G, *stable(arr(x, n)), new_x,
                          *stable(arr(g, m)))

        @guard
        def eval_jac_g(n_, x, new_x, m_, nele, irow, jcol, values, _):
            if not values:
                arr(irow, nele)[:] = self._jr
                arr(jcol, nele)[:] = self._jc
            else:
                ev.ipopt_eval(backend.JAC, *stable(arr(x, n)), new_x,
                              *stable(arr(values, nele)))

        @guard
        def eval_h(n_, x, new_x, sigma, m_, lam, new_lam, nele, irow, jcol,
                   values, _):
            if not values:
                arr(irow, nele)[:] = self._hr
                arr(jcol, nele)[:] = self._hc
            else:
                ev.ipopt_eval(backend.HESS, arr(x, n), new_x,
                              arr(values, nele), sigma, arr(lam, m))

        self._cbs = (_EVAL_F(eval_f), _EVAL_G(eval_g),
                     _EVAL_GRAD_F(eval_grad_f), _EVAL_JAC_G(eval_jac_g),
                     _EVAL_H(eval_h))
        self.last_error = None
        as_dp = lambda a: a.ctypes.data_as(_dp)     # noqa: E731
        self._nlp = lib.CreateIpoptProblem(
            n, as_dp(self.x_L), as_dp(self.x_U), m, as_dp(self.g_L),
            as_dp(self.g_U), len(jr), len(hr), 0, *self._cbs)
        if not self._nlp:
            raise RuntimeError('CreateIpoptProblem failed')

    def solve(self, dec0):
        lib, nlp = self.lib, self._nlp
        for k, v in self.str_options.items():
            lib.AddIpoptStrOption(nlp, k.encode(), str(v).encode())
        for k, v in self.num_options.items():
            lib.AddIpoptNumOption(nlp, k.encode(), v)
        for k, v in self.int_options.items():
            lib.AddIpoptIntOption(nlp, k.encode(), v)
        n, m = self.ev.n, self.ev.m
        if self.dec_scale is not None or self.constr_scale is not None \
                or self.obj_scale != 1.0:
            ds = np.ones(n) if self.dec_scale is None else self.dec_scale
            cs = np.ones(m) if self.constr_scale is None else \
                self.constr_scale
            lib.AddIpoptStrOption(nlp, b'nlp_scaling_method',
                                  b'user-scaling')
            lib.SetIpoptProblemScaling(nlp, self.obj_scale,
                                       ds.ctypes.data_as(_dp),
                                       cs.ctypes.data_as(_dp))
        x = np.array(dec0, dtype=float)
        g = np.zeros(m)
        obj = ctypes.c_double()
        mult_g, mult_L, mult_U = np.zeros(m), np.zeros(n), np.zeros(n)
        t0 = time.perf_counter()
        ev_t0 = getattr(self.ev, 'seconds', 0.0)
        try:
            status = lib.IpoptSolve(
                nlp, x.ctypes.data_as(_dp), g.ctypes.data_as(_dp),
                ctypes.cast(ctypes.byref(obj), _dp),
                mult_g.ctypes.data_as(_dp), mult_L.ctypes.data_as(_dp),
                mult_U.ctypes.data_as(_dp), None)
        finally:
            unpin = getattr(self.ev, 'unpin', None)
            if self.pin_buffers and unpin is not None:
                unpin()         # IPOPT's arrays do not outlive the solve
        total = time.perf_counter() - t0
        cb = getattr(self.ev, 'seconds', 0.0) - ev_t0
        info = {'status': int(status), 'solver': self.name, 'obj': obj.value,
                'g': g, 'mult_g': mult_g, 'mult_x_L': mult_L,
                'mult_x_U': mult_U, 'seconds_total': total,
                'seconds_callbacks': cb,
                'seconds_kkt': total - cb,     # IPOPT internal (incl. MA57)
                'last_error': self.last_error}
        return x, info

    def close(self):
        if getattr(self, '_nlp', None):
            self.lib.FreeIpoptProblem(self._nlp)
            self._nlp = None
        super().close()


def warm_up():
    """Compile the optional numba kernels now (e.g. before forking workers)."""
    D = np.stack([np.eye(2), -np.eye(2)])
    _bt_negative(D, np.zeros((1, 2, 2)))


def make_solver(problem, dec_bounds, constr_bounds, evaluator=None):
    """Solver for ``problem.ipopt(...)``: IPOPT if a libipopt is loadable
    (or ``CFEM_NLP=ipopt``), else the built-in interior-point driver."""
    evaluator = evaluator or GpuEvaluator(problem)
    choice = os.environ.get('CFEM_NLP', 'auto')
    if choice == 'ipopt' or (choice == 'auto' and find_ipopt()):
        return IpoptSolver(evaluator, dec_bounds, constr_bounds)
    return InteriorPointSolver(evaluator, dec_bounds, constr_bounds)
