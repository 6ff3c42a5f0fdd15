"""Restatement of the reference's model/problem families for the CPU oracle.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not travel to the GPU box,
so the oracle carries its own statement of the mathematics, written with sympy
matrices on top of ``oracle.engine``.  Each piece cites the reference lines it
restates; ``tests/test_oracle_reference.py`` (build container only) and the
fixtures in ``tests/golden`` check that this restatement and the reference's
own unchanged ``symfem.py`` / ``fem.py`` agree entry for entry.

Instead of the reference's cooperative multiple inheritance the families are
composed from an ordered feature tuple; the order is the reverse MRO of the
reference's classes, which is the order in which they register decisions and
functions (e.g. ``Problem(ML, ZOH)`` of attas_sp_ml_zoh.py:53 registers
innovation -> zoh -> ml).
"""

import math

import numpy as np
import sympy

from oracle import engine

#: feature tuples of the compositions the reference scripts use
KINDS = {
    'innovation': (),                       # attas_sp_innov.py
    'balanced': ('balanced',),              # attas_sp_innov_bal.py:85-87
    'ml': ('ml',),                          # attas_sp_ml.py:85-87
    'ml_zoh': ('zoh', 'ml'),                # attas_sp_ml_zoh.py:49-54
    'ndisc_zoh': ('zoh', 'ml', 'ndisc'),    # attas_sp_ml_ndisc.py:49-54
    'ml_balanced': ('balanced', 'ml'),      # mc_blackbox_cfem.py:25-30
}


def tril_pairs(n):
    """Row-major lower-triangle index pairs (symfem.py:251-252)."""
    return [(i, j) for i in range(n) for j in range(i + 1)]


def tril_diag(n):
    """Mask of the diagonal inside the tril vector (symfem.py:255-256)."""
    return np.array([i == j for i, j in tril_pairs(n)])


def tril_matrix(elems):
    """Lower-triangular sympy matrix from its tril vector (symfem.py:259-265)."""
    elems = list(np.ravel(elems))
    n = int(round((math.sqrt(8 * len(elems) + 1) - 1) / 2))
    m = sympy.zeros(n, n)
    for (i, j), e in zip(tril_pairs(n), elems):
        m[i, j] = e
    return m


def _mat(a):
    a = np.asarray(a, dtype=object)
    if a.ndim == 1:
        return sympy.Matrix(len(a), 1, list(a))
    return sympy.Matrix(a.shape[0], a.shape[1], list(a.ravel()))


def _arr(m):
    return np.array(m.tolist(), dtype=object)


def _vec(m):
    return np.array(list(m), dtype=object)


def _names(prefix, *dims):
    if len(dims) == 1:
        return [f'{prefix}{i}' for i in range(dims[0])]
    return [[f'{prefix}{i}_{j}' for j in range(dims[1])]
            for i in range(dims[0])]


def _tril_names(prefix, n):
    return [f'{prefix}{i}_{j}' for i, j in tril_pairs(n)]


class FilterErrorModel(engine.Model):
    """All model families of symfem.py, switched on by ``features``."""

    expm_order = 3          # symfem.py:178
    noise_disc_order = 1    # symfem.py:207

    def __init__(self, nx, nu, ny, features=()):
        super().__init__()
        self.nx, self.nu, self.ny = nx, nu, ny
        self.features = tuple(features)
        v = self.variables

        # symfem.py:24-43 (decision + auxiliary symbols of the predictor)
        for name, n in (('x', nx), ('en', ny), ('xnext', nx), ('xprev', nx),
                        ('enprev', ny), ('ybias', ny)):
            v[name] = _names(name, n)
        for name, shape in (('A', (nx, nx)), ('B', (nx, nu)), ('C', (ny, nx)),
                            ('D', (ny, nu)), ('Ln', (nx, ny))):
            v[name] = _names(name, *shape)
        v['sRp_tril'] = _tril_names('sRp', ny)
        self.decision |= {k for k in v if k != 'self'}
        for name, n in (('u', nu), ('y', ny), ('uprev', nu)):
            v[name] = _names(name, n)
        # symfem.py:46-48
        self.add_constraint('dynamics')
        self.add_constraint('innovation')
        self.add_objective('loglikelihood')

        for feat in self.features:
            getattr(self, '_init_' + feat)()

    # symfem.py:79-92
    def _init_balanced(self):
        v, nx, nu, ny = self.variables, self.nx, self.nu, self.ny
        v['sW_diag'] = _names('sW', nx)
        v['ctrl_orth'] = _names('ctrl_orth', nx, nx + nu)
        v['obs_orth'] = _names('obs_orth', nx, nx + ny)
        self.decision |= {'sW_diag', 'ctrl_orth', 'obs_orth'}
        for f in ('ctrl_gram', 'obs_gram', 'ctrl_orthogonality',
                  'obs_orthogonality'):
            self.add_constraint(f)

    # symfem.py:115-135
    def _init_ml(self):
        v, nx, ny = self.variables, self.nx, self.ny
        v['Kn'] = _names('Kn', nx, ny)
        v['sQ_tril'] = _tril_names('sQ', nx)
        v['sR_tril'] = _tril_names('sR', ny)
        v['sPp_tril'] = _tril_names('sPp', nx)
        v['sPc_tril'] = _tril_names('sPc', nx)
        v['pred_orth'] = _names('pred_orth', nx, 2 * nx)
        v['corr_orth'] = _names('corr_orth', nx + ny, nx + ny)
        self.decision |= {'sPp_tril', 'sPc_tril', 'sQ_tril', 'sRp_tril',
                          'sR_tril', 'Kn', 'pred_orth', 'corr_orth'}
        for f in ('pred_orthogonality', 'corr_orthogonality', 'pred_cov',
                  'corr_cov', 'kalman_gain'):
            self.add_constraint(f)

    # symfem.py:183-191
    def _init_zoh(self):
        v, nx, nu = self.variables, self.nx, self.nu
        v['dt'] = 'dt'
        v['Ac'] = _names('Ac', nx, nx)
        v['Bc'] = _names('Bc', nx, nu)
        self.decision |= {'Ac', 'Bc'}
        self.add_constraint('discretize_AB')

    # symfem.py:212-222
    def _init_ndisc(self):
        v, nx = self.variables, self.nx
        v['sQc_tril'] = _tril_names('sQc', nx)
        if 'Ac' not in v:
            v['Ac'] = _names('Ac', nx, nx)
        if 'dt' not in v:
            v['dt'] = 'dt'
        self.decision |= {'Ac', 'sQc_tril'}
        self.add_constraint('discretize_Q')

    # ---- per-sample functions (the hot path) --------------------------------
    def dynamics(self, xnext, xprev, uprev, enprev, A, B, Ln):
        """symfem.py:50-53: xnext - (A xprev + B uprev + Ln enprev)."""
        pred = _mat(A) * _mat(xprev) + _mat(B) * _mat(uprev) \
            + _mat(Ln) * _mat(enprev)
        return _vec(_mat(xnext) - pred)

    def innovation(self, y, en, x, u, C, D, ybias, sRp_tril):
        """symfem.py:55-59: y - (C x + D u + ybias) - sRp en."""
        ymodel = _mat(C) * _mat(x) + _mat(D) * _mat(u) + _mat(ybias)
        return _vec(_mat(y) - ymodel - tril_matrix(sRp_tril) * _mat(en))

    def loglikelihood(self, en, sRp_tril):
        """symfem.py:61-65: -1/2 |en|^2 - sum_i log sRp_ii."""
        sRp = tril_matrix(sRp_tril)
        logdet = sum(sympy.log(sRp[i, i]) for i in range(sRp.rows))
        return -0.5 * sum(e ** 2 for e in en) - logdet

    # ---- parameter-only functions -------------------------------------------
    def _half_gram_defect(self, orth):
        q = _mat(orth)
        resid = (q * q.T - sympy.eye(q.rows)) / 2
        # 0.5*(...) in the reference: keep the float coefficient
        resid = resid.applyfunc(lambda e: sympy.Float(1.0) * e)
        return np.array([resid[i, j] for i, j in tril_pairs(q.rows)],
                        dtype=object)

    def ctrl_gram(self, sW_diag, A, B, ctrl_orth):
        """symfem.py:94-96: sW*ctrl_orth - [A*sW, B]."""
        sW = sympy.diag(*sW_diag)
        rhs = (_mat(A) * sW).row_join(_mat(B))
        return _arr(sW * _mat(ctrl_orth) - rhs)

    def obs_gram(self, sW_diag, A, C, obs_orth):
        """symfem.py:98-100: sW*obs_orth - [A'*sW, C']."""
        sW = sympy.diag(*sW_diag)
        rhs = (_mat(A).T * sW).row_join(_mat(C).T)
        return _arr(sW * _mat(obs_orth) - rhs)

    def ctrl_orthogonality(self, ctrl_orth):
        """symfem.py:102-104."""
        return self._half_gram_defect(ctrl_orth)

    def obs_orthogonality(self, obs_orth):
        """symfem.py:106-108."""
        return self._half_gram_defect(obs_orth)

    def pred_orthogonality(self, pred_orth):
        """symfem.py:137-139."""
        return self._half_gram_defect(pred_orth)

    def corr_orthogonality(self, corr_orth):
        """symfem.py:141-143."""
        return self._half_gram_defect(corr_orth)

    def pred_cov(self, A, sPp_tril, sPc_tril, sQ_tril, pred_orth):
        """symfem.py:145-150: sPp*pred_orth - [A*sPc, sQ]."""
        rhs = (_mat(A) * tril_matrix(sPc_tril)).row_join(tril_matrix(sQ_tril))
        return _arr(tril_matrix(sPp_tril) * _mat(pred_orth) - rhs)

    def corr_cov(self, C, sR_tril, sRp_tril, sPp_tril, sPc_tril, Kn,
                 corr_orth):
        """symfem.py:152-164: [[sRp,0],[Kn,sPc]]*corr_orth - [[sR,C sPp],[0,sPp]]."""
        nx, ny = self.nx, self.ny
        sPp = tril_matrix(sPp_tril)
        left = tril_matrix(sRp_tril).row_join(sympy.zeros(ny, nx)).col_join(
            _mat(Kn).row_join(tril_matrix(sPc_tril)))
        right = tril_matrix(sR_tril).row_join(_mat(C) * sPp).col_join(
            sympy.zeros(nx, ny).row_join(sPp))
        return _arr(left * _mat(corr_orth) - right)

    def kalman_gain(self, Ln, Kn, A):
        """symfem.py:166-167: Ln - A*Kn."""
        return _arr(_mat(Ln) - _mat(A) * _mat(Kn))

    def discretize_AB(self, A, B, Ac, Bc, dt):
        """symfem.py:193-202 with the Taylor expm of symfem.py:278-289."""
        n = self.nx
        a = _mat(Ac) * dt
        term = sympy.eye(n)
        expm = sympy.eye(n)
        for i in range(1, self.expm_order + 1):
            term = (term * a) * (sympy.Float(1.0) / i)
            expm = expm + term
        bterm = dt * _mat(Bc)
        bsum = bterm
        for k in range(2, self.expm_order + 1):
            bterm = (dt / k) * _mat(Ac) * bterm
            bsum = bsum + bterm
        return _arr((_mat(A) - expm).row_join(_mat(B) - bsum))

    def discretize_Q(self, Ac, sQ_tril, sQc_tril, dt):
        """symfem.py:224-248 (Van-Loan-type series, order noise_disc_order)."""
        sQ = tril_matrix(sQ_tril)
        sQc = tril_matrix(sQc_tril)
        Q = sQ * sQ.T
        Qc = sQc * sQc.T
        acc = sympy.zeros(self.nx, self.nx)
        order = self.noise_disc_order
        for j in range(order + 1):
            for k in range(j, order + 1):
                scal = dt ** (j + k + 1) / (j + k + 1) \
                    / float(math.factorial(j)) / float(math.factorial(k))
                term = scal * (_mat(Ac) ** j) * Qc * (_mat(Ac) ** k).T
                acc += term if j == k else term + term.T
        resid = acc - Q
        return np.array([resid[i, j] for i, j in tril_pairs(self.nx)],
                        dtype=object)

    @property
    def generate_assignments(self):
        """symfem.py:67-72,169-173."""
        gen = {'nx': self.nx, 'nu': self.nu, 'ny': self.ny,
               'nty': len(self.variables['sRp_tril'])}
        if 'ml' in self.features:
            gen['ntx'] = len(self.variables['sPp_tril'])
        return gen


class FilterErrorProblem(engine.Problem):
    """All problem families of fem.py, switched on by ``features``."""

    def __init__(self, model, y, u, features=(), halo=0):
        """``halo=1`` states the left/inner shard of a time-split trajectory
        (no reference counterpart; used by the multi-process tests): ``x``
        carries one extra row owned by the right neighbour and ``dynamics``
        has N rows instead of N-1."""
        super().__init__()
        self.model = model
        self.features = tuple(features)
        self.y = np.asarray(y, dtype=float)
        self.u = np.asarray(u, dtype=float)
        N = len(self.y)
        self.N = N
        M = N - 1 + halo            # rows of the one-step functions
        self.uprev = self.u[:M]
        nx, nu, ny = model.nx, model.nu, model.ny
        assert self.y.shape == (N, ny) and self.u.shape == (N, nu) and N > 1

        # fem.py:36-44
        for name, shape in (('ybias', ny), ('sRp_tril', model.nty),
                            ('A', (nx, nx)), ('B', (nx, nu)), ('C', (ny, nx)),
                            ('D', (ny, nu)), ('Ln', (nx, ny))):
            self.add_decision(name, shape)
        x = self.add_decision('x', (N + halo, nx))
        en = self.add_decision('en', (N, ny))
        # fem.py:47-52
        D = engine.Decision
        self.add_dependent_variable('xprev', D((M, nx), x.offset))
        self.add_dependent_variable('enprev', D((M, ny), en.offset))
        self.add_dependent_variable('xnext', D((M, nx), x.offset + nx))
        if halo:
            # the per-sample functions of this shard see its own N rows of x
            self.add_dependent_variable('x', D((N, nx), x.offset))
        # fem.py:55-57
        self.add_objective(model.loglikelihood, N)
        self.add_constraint(model.dynamics, (M, nx))
        self.add_constraint(model.innovation, (N, ny))

        for feat in self.features:
            getattr(self, '_init_' + feat)()

    # fem.py:66-84
    def _init_balanced(self):
        m = self.model
        nx, nu, ny = m.nx, m.nu, m.ny
        ntx = nx * (nx + 1) // 2
        self.add_decision('sW_diag', nx)
        self.add_decision('ctrl_orth', (nx, nx + nu))
        self.add_decision('obs_orth', (nx, nx + ny))
        self.add_constraint(m.ctrl_gram, (nx, nx + nu))
        self.add_constraint(m.obs_gram, (nx, nx + ny))
        self.add_constraint(m.ctrl_orthogonality, ntx)
        self.add_constraint(m.obs_orthogonality, ntx)

    # fem.py:87-109
    def _init_ml(self):
        m = self.model
        nx, ny = m.nx, m.ny
        nxy = nx + ny
        for name in ('sPp_tril', 'sPc_tril', 'sQ_tril'):
            self.add_decision(name, m.ntx)
        self.add_decision('sR_tril', m.nty)
        self.add_decision('Kn', (nx, ny))
        self.add_decision('pred_orth', (nx, 2 * nx))
        self.add_decision('corr_orth', (nxy, nxy))
        self.add_constraint(m.pred_orthogonality, m.ntx)
        self.add_constraint(m.corr_orthogonality, nxy * (nxy + 1) // 2)
        self.add_constraint(m.pred_cov, (nx, 2 * nx))
        self.add_constraint(m.corr_cov, (nxy, nxy))
        self.add_constraint(m.kalman_gain, (nx, ny))

    # fem.py:112-128
    def _init_zoh(self):
        m = self.model
        self.add_decision('Ac', (m.nx, m.nx))
        self.add_decision('Bc', (m.nx, m.nu))
        self.add_constraint(m.discretize_AB, (m.nx, m.nx + m.nu))

    # fem.py:131-149
    def _init_ndisc(self):
        m = self.model
        self.add_decision('sQc_tril', m.ntx)
        if 'Ac' not in self.decision:
            self.add_decision('Ac', (m.nx, m.nx))
        self.add_constraint(m.discretize_Q, m.nx * (m.nx + 1) // 2)

    def variables(self, dvec):
        """fem.py:59-62,126-128,147-149."""
        out = {'y': self.y, 'u': self.u, 'uprev': self.uprev}
        if 'zoh' in self.features or 'ndisc' in self.features:
            out['dt'] = self.model.dt
        out.update(super().variables(dvec))
        return out


class TrapezoidalModel(engine.Model):
    """EXTENSION without reference counterpart: continuous-time predictor
    dx/dt = Ac x + Bc u + Lc en collocated with the trapezoidal rule.  The
    oracle of this family is this independent sympy statement (matrix form;
    the product states it with NumPy object arrays)."""

    def __init__(self, nx, nu, ny):
        super().__init__()
        self.nx, self.nu, self.ny = nx, nu, ny
        v = self.variables
        for name, n in (('x', nx), ('en', ny), ('xnext', nx), ('xprev', nx),
                        ('enprev', ny), ('ennext', ny), ('ybias', ny)):
            v[name] = _names(name, n)
        for name, shape in (('Ac', (nx, nx)), ('Bc', (nx, nu)),
                            ('C', (ny, nx)), ('D', (ny, nu)),
                            ('Lc', (nx, ny))):
            v[name] = _names(name, *shape)
        v['sRp_tril'] = _tril_names('sRp', ny)
        self.decision |= {k for k in v if k != 'self'}
        for name, n in (('u', nu), ('y', ny), ('uprev', nu), ('unext', nu)):
            v[name] = _names(name, n)
        v['dt'] = 'dt'
        self.add_constraint('trapezoid')
        self.add_constraint('innovation')
        self.add_objective('loglikelihood')

    def trapezoid(self, xnext, xprev, unext, uprev, ennext, enprev, Ac, Bc,
                  Lc, dt):
        f0 = _mat(Ac) * _mat(xprev) + _mat(Bc) * _mat(uprev) \
            + _mat(Lc) * _mat(enprev)
        f1 = _mat(Ac) * _mat(xnext) + _mat(Bc) * _mat(unext) \
            + _mat(Lc) * _mat(ennext)
        return _vec(_mat(xnext) - _mat(xprev)
                    - sympy.Float(0.5) * dt * (f0 + f1))

    innovation = FilterErrorModel.innovation
    loglikelihood = FilterErrorModel.loglikelihood

    @property
    def generate_assignments(self):
        return {'nx': self.nx, 'nu': self.nu, 'ny': self.ny,
                'nty': len(self.variables['sRp_tril'])}


class TrapezoidalProblem(engine.Problem):
    def __init__(self, model, y, u):
        super().__init__()
        self.model = model
        self.y = np.asarray(y, dtype=float)
        self.u = np.asarray(u, dtype=float)
        N = self.N = len(self.y)
        nx, nu, ny = model.nx, model.nu, model.ny
        for name, shape in (('ybias', ny), ('sRp_tril', model.nty),
                            ('Ac', (nx, nx)), ('Bc', (nx, nu)),
                            ('C', (ny, nx)), ('D', (ny, nu)),
                            ('Lc', (nx, ny))):
            self.add_decision(name, shape)
        x = self.add_decision('x', (N, nx))
        en = self.add_decision('en', (N, ny))
        D = engine.Decision
        self.add_dependent_variable('xprev', D((N - 1, nx), x.offset))
        self.add_dependent_variable('xnext', D((N - 1, nx), x.offset + nx))
        self.add_dependent_variable('enprev', D((N - 1, ny), en.offset))
        self.add_dependent_variable('ennext', D((N - 1, ny), en.offset + ny))
        self.add_objective(model.loglikelihood, N)
        self.add_constraint(model.trapezoid, (N - 1, nx))
        self.add_constraint(model.innovation, (N, ny))

    def variables(self, dvec):
        out = {'y': self.y, 'u': self.u, 'uprev': self.u[:-1],
               'unext': self.u[1:], 'dt': self.model.dt}
        out.update(super().variables(dvec))
        return out


_model_cache = {}


def make_model(kind, nx, nu, ny):
    """Compiled oracle model instance (cached: sympy differentiation is slow)."""
    key = (kind, nx, nu, ny)
    if key not in _model_cache:
        sym = TrapezoidalModel(nx, nu, ny) if kind == 'trapezoid' else \
            FilterErrorModel(nx, nu, ny, KINDS[kind])
        _model_cache[key] = sym.compile_class()
    return _model_cache[key]()


def make_problem(kind, y, u, nx, dt=None, halo=0):
    y = np.asarray(y, dtype=float)
    u = np.asarray(u, dtype=float)
    model = make_model(kind, nx, u.shape[1], y.shape[1])
    if dt is not None:
        model.dt = dt
    if kind == 'trapezoid':
        assert halo == 0
        return TrapezoidalProblem(model, y, u)
    return FilterErrorProblem(model, y, u, KINDS[kind], halo=halo)
