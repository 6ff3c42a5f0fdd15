"""Fit procedure helpers: predictor-consistent initial guesses and the
Monte-Carlo batch driver.

The reference's ``mc_blackbox_cfem.py`` is a stub: ``predict`` computes the
one-step-ahead predictor of a guess but returns nothing
(/root/reference/mc_blackbox_cfem.py:53-74) and ``estimate`` is ``pass``
(:77-78).  The procedure here completes it the way the stub points to:
run the predictor of an initial model to obtain states and innovations that
satisfy the collocation constraints exactly, normalise the innovations by
the Cholesky factor of their sample covariance, then hand the problem to the
NLP solver.
"""

import os

import numpy as np


def predictor_guess(y, u, A, B, C, D, Lun, x0=None):
    """States / normalised innovations of the predictor

        e[k]   = y[k] - C x[k] - D u[k]
        x[k+1] = A x[k] + B u[k] + Lun e[k]

    (mc_blackbox_cfem.py:53-74) as a dict of decision-variable values
    ``A B C D Ln x en sRp_tril ybias`` that satisfies ``dynamics`` and
    ``innovation`` (symfem.py:50-59) to rounding: ``sRp`` is the Cholesky
    factor of the sample innovation covariance, ``en = sRp^-1 e`` and
    ``Ln = Lun sRp``.
    """
    y, u = np.asarray(y, float), np.asarray(u, float)
    N, ny = y.shape
    nx = len(A)
    x = np.zeros((N, nx))
    if x0 is not None:
        x[0] = x0
    e = np.empty_like(y)
    for k in range(N):
        e[k] = y[k] - C @ x[k] - D @ u[k]
        if k + 1 < N:
            x[k + 1] = A @ x[k] + B @ u[k] + Lun @ e[k]
    Rp = e.T @ e / N
    sRp = np.linalg.cholesky(Rp)
    en = np.linalg.solve(sRp, e.T).T
    return {'A': A, 'B': B, 'C': C, 'D': D, 'Ln': Lun @ sRp, 'x': x,
            'en': en, 'sRp_tril': sRp[np.tril_indices(ny)],
            'ybias': np.zeros(ny)}


def _lq(M):
    """M = L Q with L lower triangular (positive diagonal), Q orthonormal rows."""
    q, r = np.linalg.qr(M.T)
    sign = np.sign(np.diag(r))
    sign[sign == 0] = 1.0
    return (r.T * sign), (q * sign).T


def kalman_guess(y, u, A, B, C, D, sQ, sR, x0=None, iters=2000, tol=1e-13):
    """Fully feasible starting point of the maximum-likelihood problem.

    Iterates the square-root (array) Riccati recursion that the reference
    states as equality constraints (/root/reference/symfem.py:145-167):

        sPp pred_orth = [A sPc, sQ]
        [[sRp, 0], [Kn, sPc]] corr_orth = [[sR, C sPp], [0, sPp]]
        Ln = A Kn

    to its fixed point by LQ factorisations, then runs the steady-state
    predictor (``predictor_guess``) so that ``dynamics`` and ``innovation``
    hold as well.  Returns decision-variable values by name.
    """
    A, C = np.asarray(A, float), np.asarray(C, float)
    nx, ny = len(A), len(C)
    sPc = np.array(sQ, dtype=float)
    for _ in range(iters):
        sPp, pred_orth = _lq(np.hstack((A @ sPc, sQ)))
        pre = np.vstack((np.hstack((sR, C @ sPp)),
                         np.hstack((np.zeros((nx, ny)), sPp))))
        post, corr_orth = _lq(pre)
        new = post[ny:, ny:]
        done = np.max(np.abs(new - sPc)) <= tol * max(1.0, np.max(np.abs(new)))
        sPc = new
        if done:
            break
    sPp, pred_orth = _lq(np.hstack((A @ sPc, sQ)))
    pre = np.vstack((np.hstack((sR, C @ sPp)),
                     np.hstack((np.zeros((nx, ny)), sPp))))
    post, corr_orth = _lq(pre)
    sRp, Kn, sPc = post[:ny, :ny], post[ny:, :ny], post[ny:, ny:]
    Ln = A @ Kn
    out = predictor_guess(y, u, A, B, C, D, Ln @ np.linalg.inv(sRp), x0)
    # the ML family ties sRp to the filter, not to the sample covariance
    e = out['en'] @ np.linalg.cholesky(
        np.atleast_2d(_tril_to_mat(out['sRp_tril'], ny))
        @ np.atleast_2d(_tril_to_mat(out['sRp_tril'], ny)).T).T
    tri_x, tri_y = np.tril_indices(nx), np.tril_indices(ny)
    out.update({
        'Ln': Ln, 'Kn': Kn, 'en': np.linalg.solve(sRp, e.T).T,
        'sRp_tril': sRp[tri_y], 'sR_tril': np.asarray(sR)[tri_y],
        'sQ_tril': np.asarray(sQ)[tri_x], 'sPp_tril': sPp[tri_x],
        'sPc_tril': sPc[tri_x], 'pred_orth': pred_orth,
        'corr_orth': corr_orth})
    return out


def _tril_to_mat(elem, n):
    m = np.zeros((n, n))
    m[np.tril_indices(n)] = elem
    return m


# ----------------------------------------------------------------------------
# problem set-up in the manner of the reference scripts
# ----------------------------------------------------------------------------

def ml_setup(problem, fix=None):
    """Bounds and scaling of a maximum-likelihood problem as in
    /root/reference/attas_sp_ml.py:113-151: square-root covariance diagonals
    bounded below, ``sR`` diagonal, innovation / covariance / gain constraints
    scaled by 100 and the square-root factors and gains by 100.  ``fix`` maps
    variable names to values held fixed through equal bounds (the scripts fix
    ``C`` and ``D`` that way).

    Returns ``(dec_bounds, constr_bounds, (obj_scale, dec_scale, constr_scale))``.
    """
    from . import models
    dec_bounds = np.repeat([[-np.inf], [np.inf]], problem.ndec, axis=-1)
    lo, hi = problem.variables(dec_bounds[0]), problem.variables(dec_bounds[1])
    for name, value in (fix or {}).items():
        lo[name][...] = value
        hi[name][...] = value

    def diag(name):
        n = models.tril_mat(np.zeros(problem.decision[name].size)).shape[0]
        return models.tril_diag(n)
    lo['sRp_tril'][diag('sRp_tril')] = 1e-6
    for name in ('sPp_tril', 'sPc_tril', 'sQ_tril'):
        if name in lo:
            lo[name][diag(name)] = 0.0
    if 'sR_tril' in lo:
        d = diag('sR_tril')
        lo['sR_tril'][d] = 1e-6
        lo['sR_tril'][~d] = 0.0
        hi['sR_tril'][~d] = 0.0
    if 'sW_diag' in lo:
        lo['sW_diag'][...] = 0.0
    constr_bounds = np.zeros((2, problem.ncons))
    constr_scale = np.ones(problem.ncons)
    cs = problem.unpack_constraints(constr_scale)
    for name in ('innovation', 'pred_cov', 'corr_cov', 'kalman_gain'):
        if name in cs:
            cs[name][...] = 100.0
    dec_scale = np.ones(problem.ndec)
    ds = problem.variables(dec_scale)
    for name in ('Ln', 'sRp_tril', 'sPp_tril', 'sPc_tril', 'sQ_tril',
                 'sR_tril', 'Kn'):
        if name in ds:
            ds[name][...] = 100.0
    return dec_bounds, constr_bounds, (-1.0, dec_scale, constr_scale)


def start_point(problem, guess):
    """Decision vector from a dict of variable values (others zero)."""
    dec0 = np.zeros(problem.ndec)
    var0 = problem.variables(dec0)
    for name, value in guess.items():
        if name in problem.decision:
            var0[name][...] = value
    return dec0


def solve(problem, dec0, dec_bounds, constr_bounds, scaling, tol=1e-9,
          max_iter=500, **options):
    """``problem.ipopt(...)`` exactly as the reference scripts call it
    (attas_sp_ml.py:153-159)."""
    with problem.ipopt(dec_bounds, constr_bounds) as nlp:
        nlp.add_str_option('linear_solver', 'ma57')
        nlp.add_num_option('tol', tol)
        nlp.add_int_option('max_iter', max_iter)
        for key, value in options.items():
            nlp.add_num_option(key, value)
        nlp.set_scaling(*scaling)
        return nlp.solve(dec0)


# ----------------------------------------------------------------------------
# Monte-Carlo batches: whole problems per GPU, evaluated together
# ----------------------------------------------------------------------------

class BatchFitter:
    """Fit many same-shaped problems on one GPU in lock-step.

    The reference's Monte-Carlo loop is serial (``for datafile in datafiles``,
    /root/reference/mc_blackbox_cfem.py:115) and its body is a stub.  Here
    every problem of the batch runs its own interior-point iteration
    (``nlp.InteriorPointSolver.solve_steps``) while the callbacks of ALL
    problems of a round are served by one launch of the fused kernel with
    ``blockIdx.y`` = problem (``cfem_create(batch=B)``); the KKT
    factorisations stay on the host (thread pool).  No collective is needed:
    with several GPUs each rank takes a slice of the experiments.
    """

    def __init__(self, problems, device=0, threads=None):
        from . import backend, nlp
        self.problems = problems
        p0 = problems[0]
        st = p0.structure
        key = st.key()
        for p in problems[1:]:
            if p.structure.key() != key or p.structure.N != st.N:
                raise ValueError('a batch needs same-shaped problems')
        self.B = B = len(problems)
        self.lib = backend.Library.for_structure(st)
        data = [np.stack([np.ascontiguousarray(p.structure.data[i]['source'],
                                               dtype=float)
                          for p in problems])
                for i in range(len(st.data))]
        self.handle = backend.Handle(self.lib, st.N, data, st.scalar_values,
                                     batch=B, device=device)
        self.buf = backend.HostBuffers(self.handle)
        self.threads = threads or min(B, os.cpu_count() or 1)
        self.launches = 0
        self.seconds_gpu = 0.0
        self._nlp = nlp
        self._backend = backend

    def fit(self, dec0s, dec_bounds, constr_bounds, scaling, tol=1e-8,
            max_iter=300, options=None):
        """Returns ``[(decopt, info)] * B``; every argument but ``dec0s`` may
        be a single value shared by the batch or a list of B.  ``options``:
        numeric solver options (``mu_init``, ``bound_push`` ...)."""
        import concurrent.futures as cf
        import time
        nlp, backend = self._nlp, self._backend
        B, h, buf = self.B, self.handle, self.buf
        ndec, ncons = h.ndec, h.ncons

        def per(i, arg):
            return arg[i] if isinstance(arg, list) else arg

        solvers, steps = [], []
        for i, p in enumerate(self.problems):
            ev = _StructureOnly(p)
            s = nlp.InteriorPointSolver(ev, per(i, dec_bounds),
                                        per(i, constr_bounds))
            s.add_num_option('tol', tol)
            s.add_int_option('max_iter', max_iter)
            for key, value in (options or {}).items():
                s.add_num_option(key, value)
            s.set_scaling(*per(i, scaling))
            solvers.append(s)
            steps.append(s.solve_steps(dec0s[i]))
        sigma = _common_obj_scale([s.obj_scale for s in solvers])
        requests = [next(g) for g in steps]
        results = [None] * B
        active = set(range(B))
        dv = buf.dvec.reshape(B, ndec)
        lv = buf.lam.reshape(B, ncons)
        dv[...] = np.asarray(dec0s)
        lv[...] = 0.0
        views = {k: getattr(buf, k).reshape(B, -1)
                 for k in ('grad', 'g', 'jac', 'hess')}

        def advance(i):
            req = requests[i]
            if req[0] == 'all':
                res = (float(buf.f[i]), views['grad'][i].copy(),
                       views['g'][i].copy(), views['jac'][i].copy(),
                       views['hess'][i].copy())
            else:
                res = (float(buf.f[i]), views['g'][i].copy())
            try:
                return steps[i].send(res)
            except StopIteration as stop:
                results[i] = stop.value
                return None

        with cf.ThreadPoolExecutor(self.threads) as pool:
            while active:
                want_all = False
                for i in active:
                    req = requests[i]
                    dv[i] = req[1]
                    if req[0] == 'all':
                        lv[i] = req[3]
                        want_all = True
                t0 = time.perf_counter()
                if want_all:
                    buf.upload(sigma)
                    h.eval(backend.ALL)
                    buf.fetch_all()
                else:
                    buf.upload()
                    h.eval(backend.F | backend.G)
                    h.fetch_async(backend.F, buf.f)
                    h.fetch_async(backend.G, buf.g)
                    h.synchronize()
                self.seconds_gpu += time.perf_counter() - t0
                self.launches += 1
                order = sorted(active)
                for i, nxt in zip(order, pool.map(advance, order)):
                    if nxt is None:
                        active.discard(i)
                    else:
                        requests[i] = nxt
        return results

    def close(self):
        self.buf.close()
        self.handle.close()


def _common_obj_scale(scales):
    """The batched kernel takes ONE obj_factor per launch (a kernel argument,
    like IPOPT's eval_h): every problem of a batch must use the same
    ``obj_scale``."""
    scales = [float(v) for v in scales]
    if any(v != scales[0] for v in scales[1:]):
        raise ValueError('all problems of a batch must share obj_scale (the '
                         'batched launch has a single obj_factor); got '
                         f'{sorted(set(scales))}')
    return scales[0]


class _StructureOnly:
    """Evaluator stub for solvers driven through ``solve_steps``: sizes and
    sparsity structure only (the batch driver serves the evaluations)."""

    def __init__(self, problem):
        from . import nlp
        self.n, self.m = problem.ndec, problem.ncons
        self._p = problem
        self._ts = nlp.GpuEvaluator.time_structure
        self.problem = problem

    def jac_structure(self):
        return self._p.constr_jac_ind()

    def hess_structure(self):
        return self._p.lag_hess_ind()

    def time_structure(self):
        return self._ts(self)


def balanced_guess(A, B, C):
    """Balanced realisation of (A, B, C) and the square-root Gramian variables
    of the reference's balancing constraints (/root/reference/symfem.py:94-108):

        sW ctrl_orth = [A sW, B],   sW obs_orth = [A' sW, C'],

    rows of ``ctrl_orth`` / ``obs_orth`` orthonormal, i.e. the controllability
    and observability Gramians both equal ``diag(sW_diag)**2``.  Returns
    ``(T, values)`` with ``x_balanced = T^-1 x`` and the decision-variable
    values ``A B C sW_diag ctrl_orth obs_orth`` in balanced coordinates.
    """
    import scipy.linalg as sla
    A, B, C = (np.asarray(m, float) for m in (A, B, C))
    Wc = sla.solve_discrete_lyapunov(A, B @ B.T)
    Wo = sla.solve_discrete_lyapunov(A.T, C.T @ C)
    Lc = np.linalg.cholesky(Wc)
    Lo = np.linalg.cholesky(Wo)
    U, sv, Vt = np.linalg.svd(Lo.T @ Lc)
    T = Lc @ Vt.T / np.sqrt(sv)             # balancing transformation
    Ti = (U / np.sqrt(sv)).T @ Lo.T
    Ab, Bb, Cb = Ti @ A @ T, Ti @ B, C @ T
    sW = np.sqrt(sv)
    ctrl_orth = np.hstack((Ab * sW, Bb)) / sW[:, None]
    obs_orth = np.hstack((Ab.T * sW, Cb.T)) / sW[:, None]
    return T, {'A': Ab, 'B': Bb, 'C': Cb, 'sW_diag': sW,
               'ctrl_orth': ctrl_orth, 'obs_orth': obs_orth}


class ParallelBatchFitter:
    """``BatchFitter`` with the host work spread over worker PROCESSES.

    The interior-point iterations (KKT factorisations, line-search logic) are
    host work that Python threads cannot overlap; here ``workers`` processes
    each advance a slice of the batch, while the parent -- the only process
    that talks to the GPU -- serves every round's evaluation requests with one
    batched launch.  Requests and results travel through anonymous shared
    memory that the parent page-locks (``cfem_host_register``), so the DMA
    engines read the workers' decision vectors and write the results in place.

    Round protocol (two barriers):
      workers  write x (and lambda) of their active problems + request kind
      ---- barrier ----
      parent   H2D, one fused launch for the whole batch, D2H
      ---- barrier ----
      workers  read their results, advance their solvers (in parallel)
    """

    def __init__(self, problems, device=0, workers=None):
        import mmap
        import multiprocessing as mp
        from . import backend
        self.problems = problems
        p0 = problems[0]
        st = p0.structure
        key = st.key()
        for p in problems[1:]:
            if p.structure.key() != key or p.structure.N != st.N:
                raise ValueError('a batch needs same-shaped problems')
        from . import nlp
        nlp.warm_up()               # JIT-compile before forking
        for p in problems:          # index arrays: build once, before forking
            p.constr_jac_ind()
            p.lag_hess_ind()
        self.B = B = len(problems)
        self.workers = W = max(1, min(workers or (os.cpu_count() or 2) - 1, B))
        self.sizes = sizes = {'dvec': p0.ndec, 'lam': p0.ncons, 'f': 1,
                              'grad': p0.ndec, 'g': p0.ncons,
                              'jac': p0.nnzjac, 'hess': p0.nnzhess}
        self._maps, self.sh = {}, {}
        for name, n in sizes.items():
            nbytes = max(8, B * n * 8)
            m = mmap.mmap(-1, nbytes)       # MAP_SHARED | MAP_ANONYMOUS
            self._maps[name] = m
            self.sh[name] = np.frombuffer(m, dtype=np.float64,
                                          count=B * n).reshape(B, n)
        mk = mmap.mmap(-1, max(8, B * 8))
        self._maps['kind'] = mk
        self.kind = np.frombuffer(mk, dtype=np.int64, count=B)
        self.ctx = mp.get_context('fork')
        self.barrier = self.ctx.Barrier(W + 1)
        self.results = self.ctx.Queue()
        self.device = device
        self._st = st
        self._backend = backend
        self.launches = 0
        self.seconds_gpu = 0.0

    def _worker(self, mine, dec0s, dec_bounds, constr_bounds, scaling, tol,
                max_iter, options=None):
        import threading
        import traceback
        try:
            self._worker_body(mine, dec0s, dec_bounds, constr_bounds, scaling,
                              tol, max_iter, options)
        except threading.BrokenBarrierError:
            self.results.put(('error', f'worker {mine[:1]}: barrier broken '
                              'by another process'))
        except BaseException:       # noqa: BLE001 -- must not leave the others waiting
            self.barrier.abort()
            self.results.put(('error', traceback.format_exc()))

    def _worker_body(self, mine, dec0s, dec_bounds, constr_bounds, scaling,
                     tol, max_iter, options=None):
        from . import nlp
        try:                        # one BLAS thread per worker process
            import threadpoolctl
            threadpoolctl.threadpool_limits(1)
        except Exception:
            pass
        sh, kind = self.sh, self.kind
        wait = lambda: self.barrier.wait(self.BARRIER_TIMEOUT_S)  # noqa: E731
        steps, requests, active = {}, {}, set(mine)

        def per(i, arg):
            return arg[i] if isinstance(arg, list) else arg
        for i in mine:
            s = nlp.InteriorPointSolver(_StructureOnly(self.problems[i]),
                                        per(i, dec_bounds),
                                        per(i, constr_bounds))
            s.add_num_option('tol', tol)
            s.add_int_option('max_iter', max_iter)
            for key, value in (options or {}).items():
                s.add_num_option(key, value)
            s.set_scaling(*per(i, scaling))
            steps[i] = s.solve_steps(dec0s[i])
            requests[i] = next(steps[i])
        done = []
        while True:
            for i in mine:
                if i in active:
                    req = requests[i]
                    sh['dvec'][i] = req[1]
                    if req[0] == 'all':
                        sh['lam'][i] = req[3]
                        kind[i] = 2
                    else:
                        kind[i] = 1
                else:
                    kind[i] = 0
            wait()                      # requests are posted
            wait()                      # results are in shared memory
            if kind[0] < 0:             # parent: everybody is done
                break
            for i in list(active):
                if requests[i][0] == 'all':
                    res = (float(sh['f'][i, 0]), sh['grad'][i].copy(),
                           sh['g'][i].copy(), sh['jac'][i].copy(),
                           sh['hess'][i].copy())
                else:
                    res = (float(sh['f'][i, 0]), sh['g'][i].copy())
                try:
                    requests[i] = steps[i].send(res)
                except StopIteration as stop:
                    x, info = stop.value
                    done.append((i, x, {k: v for k, v in info.items()
                                        if not isinstance(v, np.ndarray)}))
                    active.discard(i)
        self.results.put(('done', done))

    #: a round is one batched launch (parent) or one interior-point iteration
    #: of a worker's slice; a side that does not show up within this time has
    #: died or hung
    BARRIER_TIMEOUT_S = 900.0

    def fit(self, dec0s, dec_bounds, constr_bounds, scaling, tol=1e-8,
            max_iter=300, options=None):
        import threading
        import time
        backend = self._backend
        B, W = self.B, self.workers
        slices = [list(range(w, B, W)) for w in range(W)]
        procs = [self.ctx.Process(target=self._worker,
                                  args=(sl, dec0s, dec_bounds, constr_bounds,
                                        scaling, tol, max_iter, options),
                                  daemon=True)
                 for sl in slices]
        for pr in procs:            # fork BEFORE this process touches CUDA
            pr.start()
        st = self._st
        try:
            lib = backend.Library.for_structure(st)
            data = [np.stack([np.ascontiguousarray(
                p.structure.data[i]['source'], dtype=float)
                for p in self.problems]) for i in range(len(st.data))]
            h = backend.Handle(lib, st.N, data, st.scalar_values, batch=B,
                               device=self.device)
        except BaseException:
            self.barrier.abort()        # the workers wait at the barrier
            for pr in procs:
                pr.join(30)
            raise
        registered = []
        for name, arr in self.sh.items():
            if lib.cfem_host_register(arr.ctypes.data, arr.nbytes) == 0:
                registered.append(arr)
        flat = {k: v.reshape(-1) for k, v in self.sh.items()}
        flat['lam'][:] = 0.0
        wait = lambda: self.barrier.wait(self.BARRIER_TIMEOUT_S)  # noqa: E731
        failure = None
        try:
            sigma = _common_obj_scale(
                [sc[0] for sc in scaling] if isinstance(scaling, list)
                else [scaling[0]])
            while True:
                wait()
                kinds = self.kind.copy()
                if not kinds.any():
                    self.kind[0] = -1
                    wait()
                    break
                t0 = time.perf_counter()
                h.set_dvec(flat['dvec'])
                if (kinds == 2).any():
                    h.set_multipliers(sigma, flat['lam'])
                    h.eval(backend.ALL)
                    for bit, name in ((backend.F, 'f'), (backend.GRAD, 'grad'),
                                      (backend.G, 'g'), (backend.JAC, 'jac'),
                                      (backend.HESS, 'hess')):
                        h.fetch_async(bit, flat[name])
                else:
                    h.eval(backend.F | backend.G)
                    h.fetch_async(backend.F, flat['f'])
                    h.fetch_async(backend.G, flat['g'])
                h.synchronize()
                self.seconds_gpu += time.perf_counter() - t0
                self.launches += 1
                wait()
        except threading.BrokenBarrierError:
            failure = 'a worker process failed or timed out'
        except BaseException:
            self.barrier.abort()        # release the workers, then re-raise
            raise
        finally:
            for arr in registered:
                lib.cfem_host_unregister(arr.ctypes.data)
            h.close()
        out = [None] * B
        errors = []
        import queue
        for _ in procs:
            try:
                tag, payload = self.results.get(timeout=60 if failure else 600)
            except queue.Empty:
                errors.append('a worker did not report')
                continue
            if tag == 'error':
                errors.append(payload)
                continue
            for i, x, info in payload:
                out[i] = (x, info)
        for pr in procs:
            pr.join(60)
            if pr.is_alive():
                pr.kill()
        if failure or errors:
            raise RuntimeError('ParallelBatchFitter: ' + (failure or '')
                               + '\n' + '\n'.join(errors))
        return out


# ----------------------------------------------------------------------------
# Monte-Carlo fit procedure: one attempt after the other for what did not solve
# ----------------------------------------------------------------------------

#: Attempts of the Monte-Carlo fit procedure, in order.  The reference leaves
#: the procedure open (``estimate`` is ``pass``, mc_blackbox_cfem.py:77-78).
#: Every attempt starts from the same predictor-consistent point
#: (``kalman_guess``: the ``predict`` step of mc_blackbox_cfem.py:53-74 plus the
#: filter variables); what changes is how the interior-point iteration treats
#: the sign bounds of the square-root covariance factors, which is where the
#: single-attempt runs of round 1 stalled (barrier parameter far ahead of the
#: constraint violation, multipliers of the bounds blowing up):
#:   1  IPOPT's defaults (mu_init 0.1, bound_push 1e-2);
#:   2  a smaller initial barrier parameter;
#:   3  a larger push away from the bounds;
#:   4  without the sign bounds on sPp / sPc / sQ / sW (the factors are only
#:      determined up to the sign of their columns; sRp and sR keep theirs).
#:   5  staged, as the reference's stub points to (an innovation-form problem
#:      first, mc_blackbox_cfem.py:117-118): fit the BalancedDT problem --
#:      InnovationDT plus the balancing constraints, which remove the
#:      similarity-transform gauge that makes the bare InnovationDT problem
#:      singular -- then start ML+Balanced from its solution with
#:      Q = K Rp K' and R = Rp read off the fitted innovation form
#:      (``staged_start``).
MC_ATTEMPTS = (
    {'name': 'default', 'options': {}},
    {'name': 'mu_init=1e-2', 'options': {'mu_init': 1e-2}},
    {'name': 'bound_push=1e-1', 'options': {'bound_push': 1e-1}},
    {'name': 'free factor signs', 'options': {}, 'free_factor_signs': True},
    {'name': 'staged: BalancedDT fit, then ML+Balanced from it',
     'options': {}, 'staged': True},
)


def balanced_setup(problem):
    """Bounds and scaling of a BalancedDT problem in the manner of
    /root/reference/attas_sp_innov_bal.py and blackbox_innov_bal.py:76-96."""
    from . import models
    dec_bounds = np.repeat([[-np.inf], [np.inf]], problem.ndec, axis=-1)
    lo = problem.variables(dec_bounds[0])
    ny = problem.model.ny
    lo['sRp_tril'][models.tril_diag(ny)] = 1e-6
    lo['sW_diag'][...] = 0.0
    constr_scale = np.ones(problem.ncons)
    problem.unpack_constraints(constr_scale)['innovation'][...] = 100.0
    dec_scale = np.ones(problem.ndec)
    ds = problem.variables(dec_scale)
    ds['Ln'][...] = 100.0
    ds['sRp_tril'][...] = 100.0
    return dec_bounds, np.zeros((2, problem.ncons)), (-1.0, dec_scale,
                                                      constr_scale)


def staged_start(ml_problem, bal_problem, bal_dec):
    """Feasible start of the ML+Balanced problem from a BalancedDT solution:
    re-balance (A, B, C), read Q = K Rp K' and R = diag(Rp) off the fitted
    innovation form (K = Ln sRp^-1, Rp = sRp sRp'), iterate the square-root
    Riccati recursion for them and run the predictor (``kalman_guess``)."""
    from . import models
    v = bal_problem.variables(np.asarray(bal_dec))
    A, B, C, D, Ln = (np.array(v[k]) for k in ('A', 'B', 'C', 'D', 'Ln'))
    nx = len(A)
    sRp = models.tril_mat(v['sRp_tril'])
    Rp = sRp @ sRp.T
    K = Ln @ np.linalg.inv(sRp)
    T, bal = balanced_guess(A, B, C)
    Ti = np.linalg.inv(T)
    Kb = Ti @ K
    sQ = np.linalg.cholesky(Kb @ Rp @ Kb.T + 1e-6 * np.eye(nx))
    sR = np.diag(np.sqrt(np.diag(Rp)))
    g = kalman_guess(bal_problem.y, bal_problem.u, bal['A'], bal['B'],
                     bal['C'], D, sQ, sR, x0=Ti @ v['x'][0])
    g.update({k: bal[k] for k in ('sW_diag', 'ctrl_orth', 'obs_orth')})
    g['ybias'] = np.array(v['ybias'])
    return start_point(ml_problem, g)


def staged_attempt(make_fitter, problems, dec0s, dec_bounds, constr_bounds,
                   scaling, tol, max_iter):
    """Attempt 5 of ``MC_ATTEMPTS`` for a list of ML+Balanced problems:
    two batches, BalancedDT first.  Returns ``([(decopt, info)], fitters)``."""
    from . import families
    bal = [families.make_problem('balanced', p.y, p.u, p.model.nx)
           for p in problems]
    bal0 = []
    for p, q, d in zip(problems, bal, dec0s):
        have = p.variables(np.asarray(d))
        bal0.append(start_point(q, {n: have[n] for n in q.decision}))
    db, cb, sc = balanced_setup(bal[0])
    f1 = make_fitter(bal)
    try:
        out1 = f1.fit(bal0, db, cb, sc, tol=tol, max_iter=max_iter)
    finally:
        if getattr(f1, 'close', None):
            f1.close()
    starts = []
    for p, q, d, (x1, _) in zip(problems, bal, dec0s, out1):
        try:
            starts.append(staged_start(p, q, x1))
        except (np.linalg.LinAlgError, ValueError):
            starts.append(np.asarray(d))        # keep the original start
    f2 = make_fitter(problems)
    try:
        out2 = f2.fit(starts, dec_bounds, constr_bounds, scaling, tol=tol,
                      max_iter=max_iter)
    finally:
        if getattr(f2, 'close', None):
            f2.close()
    for (_, i1), (_, i2) in zip(out1, out2):
        i2['stage1_status'] = i1['status']
        i2['stage1_iterations'] = i1['iterations']
        for key in ('iterations', 'seconds_kkt', 'callback_calls',
                    'seconds_callbacks', 'seconds_total'):
            if key in i1 and key in i2:
                i2[key] = i2[key] + i1[key]
    return out2, (f1, f2)


def free_factor_signs(problem, dec_bounds):
    """Copy of ``dec_bounds`` without the lower bounds on the diagonals of
    ``sPp`` / ``sPc`` / ``sQ`` and on ``sW_diag`` (``ml_setup``)."""
    out = np.array(dec_bounds, dtype=float)
    lo = problem.variables(out[0])
    for name in ('sPp_tril', 'sPc_tril', 'sQ_tril', 'sW_diag'):
        if name in lo:
            lo[name][...] = -np.inf
    return out


def solved(info):
    return info is not None and str(info['status']).startswith('solved')


def fit_with_retries(make_fitter, problems, dec0s, dec_bounds, constr_bounds,
                     scaling, tol=1e-6, max_iter=400, attempts=MC_ATTEMPTS,
                     log=None):
    """Run ``attempts`` one after the other on the problems that have not
    solved yet; every attempt is ONE batch (``make_fitter(subset)`` returns a
    ``BatchFitter`` / ``ParallelBatchFitter`` for a list of problems).

    Returns ``(results, report)``: ``results[i] = (decopt, info)`` of the
    attempt that solved problem i (or of the last one), ``info['attempt']`` its
    index; ``report`` lists per attempt the problems tried / solved, the wall
    time, the batched launch rounds and the GPU callback seconds.
    """
    import time
    n = len(problems)
    results = [None] * n
    todo = list(range(n))
    report = []
    for k, att in enumerate(attempts):
        if not todo:
            break
        sub = [problems[i] for i in todo]
        bounds = dec_bounds
        if att.get('free_factor_signs'):
            bounds = free_factor_signs(sub[0], dec_bounds)
        t0 = time.perf_counter()
        if att.get('staged'):
            out, fitters = staged_attempt(
                make_fitter, sub, [dec0s[i] for i in todo], bounds,
                constr_bounds, scaling, tol, max_iter)
        else:
            fitter = make_fitter(sub)
            fitters = (fitter,)
            try:
                out = fitter.fit([dec0s[i] for i in todo], bounds,
                                 constr_bounds, scaling, tol=tol,
                                 max_iter=max_iter,
                                 options=att.get('options'))
            finally:
                close = getattr(fitter, 'close', None)
                if close:
                    close()
        wall = time.perf_counter() - t0
        still = []
        for i, res in zip(todo, out):
            res[1]['attempt'] = k
            if solved(res[1]) or results[i] is None or k == len(attempts) - 1:
                results[i] = res
            if not solved(res[1]):
                still.append(i)
        rec = {'attempt': att['name'], 'tried': len(todo),
               'solved': len(todo) - len(still), 'wall_s': wall,
               'launch_rounds': sum(getattr(f, 'launches', 0) or 0
                                    for f in fitters),
               'seconds_gpu_callbacks': sum(getattr(f, 'seconds_gpu', 0.0)
                                            or 0.0 for f in fitters)}
        report.append(rec)
        if log:
            log(rec)
        todo = still
    return results, report
