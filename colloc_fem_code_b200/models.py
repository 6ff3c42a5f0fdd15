"""Filter-error model families (host-side mirror of /root/reference/symfem.py).

Same class names, constructor signature ``(nx, nu, ny)``, variable names,
registration order and cooperative multiple inheritance as the reference, so a
script written against ``symfem`` works with ``colloc_fem_code_b200.models``
(``colloc_fem_code_b200.compat`` also installs it under the name ``symfem``).
The expressions are only *declared* here; they are evaluated by the generated
CUDA kernels.

Steady-state innovation-form predictor with normalised innovations:

    x[k+1] = A x[k] + B u[k] + Ln en[k]                   (dynamics defect)
    y[k]   = C x[k] + D u[k] + ybias + sRp en[k]          (innovation)
    l[k]   = -1/2 |en[k]|^2 - log det sRp                 (log-likelihood)
"""

import math

import numpy as np
import sympy

from . import symoptim


# ---------------------------------------------------------------------------
# lower-triangle helpers (script-facing API, symfem.py:251-265)
# ---------------------------------------------------------------------------

def tril_ind(n):
    """Row-major index pairs of the lower triangle, == ``np.tril_indices``."""
    for i in range(n):
        for j in range(i + 1):
            yield (i, j)


def tril_diag(n):
    """Boolean mask selecting the diagonal entries of a tril vector."""
    return np.fromiter((i == j for i, j in tril_ind(n)), dtype=bool)


def tril_mat(elem, *legacy):
    """Lower-triangular matrix from its tril vector.

    ``tril_mat(elem)`` is the reference signature (symfem.py:259).  The stale
    scripts call ``tril_mat(n, elem)`` (hfb320_sqrt_zoh.py:196,
    blackbox_innov_bal.py:116); that form is accepted too.
    """
    if legacy:
        elem = legacy[0]
    elem = np.asarray(elem)
    n = int(round((math.sqrt(8 * len(elem) + 1) - 1) / 2))
    mat = np.zeros((n, n), dtype=elem.dtype)
    rows, cols = np.tril_indices(n)
    mat[rows, cols] = elem
    return mat


def expm_taylor(a, order):
    """Truncated Taylor series of the matrix exponential (symfem.py:278-289)."""
    a = np.asarray(a)
    assert a.ndim == 2 and a.shape[0] == a.shape[1]
    term = series = np.eye(a.shape[0])
    for i in range(1, order + 1):
        term = (term @ a) / i
        series = series + term
    return series


def _vector(prefix, n):
    return [f'{prefix}{i}' for i in range(n)]


def _matrix(prefix, rows, cols):
    return [[f'{prefix}{i}_{j}' for j in range(cols)] for i in range(rows)]


def _tril(prefix, n):
    return [f'{prefix}{i}_{j}' for i, j in tril_ind(n)]


def _half_gram_defect(q):
    """tril of (q q' - I)/2: orthonormal-rows constraint."""
    resid = 0.5 * (q @ q.T - np.eye(len(q)))
    return [resid[ij] for ij in tril_ind(len(q))]


# ---------------------------------------------------------------------------
# model families
# ---------------------------------------------------------------------------

class InnovationDTModel(symoptim.Model):
    """Discrete-time innovation-form predictor (symfem.py:11-72)."""

    def __init__(self, nx, nu, ny):
        super().__init__()
        self.nx, self.nu, self.ny = nx, nu, ny

        v = self.variables
        # per-sample decision variables and their shifted views
        v['x'] = _vector('x', nx)
        v['en'] = _vector('en', ny)
        v['xnext'] = _vector('xnext', nx)
        v['xprev'] = _vector('xprev', nx)
        v['enprev'] = _vector('enprev', ny)
        # parameters
        v['ybias'] = _vector('ybias', ny)
        v['A'] = _matrix('A', nx, nx)
        v['B'] = _matrix('B', nx, nu)
        v['C'] = _matrix('C', ny, nx)
        v['D'] = _matrix('D', ny, nu)
        v['Ln'] = _matrix('Ln', nx, ny)
        v['sRp_tril'] = _tril('sRp', ny)
        self.decision.update(name for name in v if name != 'self')
        # data
        v['u'] = _vector('u', nu)
        v['y'] = _vector('y', ny)
        v['uprev'] = _vector('uprev', nu)

        self.add_constraint('dynamics')
        self.add_constraint('innovation')
        self.add_objective('loglikelihood')

    def dynamics(self, xnext, xprev, uprev, enprev, A, B, Ln):
        """One-step predictor defect (symfem.py:50-53)."""
        return xnext - (A @ xprev + B @ uprev + Ln @ enprev)

    def innovation(self, y, en, x, u, C, D, ybias, sRp_tril):
        """Output equation with normalised innovations (symfem.py:55-59)."""
        return y - (C @ x + D @ u + ybias) - tril_mat(sRp_tril) @ en

    def loglikelihood(self, en, sRp_tril):
        """Per-sample log-likelihood term (symfem.py:61-65)."""
        logdet = sum(sympy.log(d) for d in tril_mat(sRp_tril).diagonal())
        return -0.5 * (en ** 2).sum() - logdet

    @property
    def generate_assignments(self):
        own = {'nx': self.nx, 'nu': self.nu, 'ny': self.ny,
               'nty': len(self.variables['sRp_tril'])}
        return {**own, **getattr(super(), 'generate_assignments', {})}


class BalancedDTModel(InnovationDTModel):
    """Adds the balanced-realisation constraints (symfem.py:75-108)."""

    def __init__(self, nx, nu, ny):
        super().__init__(nx, nu, ny)
        v = self.variables
        v['sW_diag'] = _vector('sW', nx)
        v['ctrl_orth'] = _matrix('ctrl_orth', nx, nx + nu)
        v['obs_orth'] = _matrix('obs_orth', nx, nx + ny)
        self.decision.update(('sW_diag', 'ctrl_orth', 'obs_orth'))
        for name in ('ctrl_gram', 'obs_gram', 'ctrl_orthogonality',
                     'obs_orthogonality'):
            self.add_constraint(name)

    def ctrl_gram(self, sW_diag, A, B, ctrl_orth):
        """sW ctrl_orth = [A sW, B] (square-root controllability Gramian)."""
        return sW_diag[:, None] * ctrl_orth - np.hstack((A * sW_diag, B))

    def obs_gram(self, sW_diag, A, C, obs_orth):
        """sW obs_orth = [A' sW, C'] (square-root observability Gramian)."""
        return sW_diag[:, None] * obs_orth - np.hstack((A.T * sW_diag, C.T))

    def ctrl_orthogonality(self, ctrl_orth):
        return _half_gram_defect(ctrl_orth)

    def obs_orthogonality(self, obs_orth):
        return _half_gram_defect(obs_orth)


class MaximumLikelihoodDTModel(InnovationDTModel):
    """Adds the square-root Riccati constraints (symfem.py:111-173)."""

    def __init__(self, nx, nu, ny):
        super().__init__(nx, nu, ny)
        v = self.variables
        v['Kn'] = _matrix('Kn', nx, ny)
        v['sQ_tril'] = _tril('sQ', nx)
        v['sR_tril'] = _tril('sR', ny)
        v['sPp_tril'] = _tril('sPp', nx)
        v['sPc_tril'] = _tril('sPc', nx)
        v['pred_orth'] = _matrix('pred_orth', nx, 2 * nx)
        v['corr_orth'] = _matrix('corr_orth', nx + ny, nx + ny)
        self.decision.update(('sPp_tril', 'sPc_tril', 'sQ_tril', 'sRp_tril',
                              'sR_tril', 'Kn', 'pred_orth', 'corr_orth'))
        for name in ('pred_orthogonality', 'corr_orthogonality', 'pred_cov',
                     'corr_cov', 'kalman_gain'):
            self.add_constraint(name)

    def pred_orthogonality(self, pred_orth):
        return _half_gram_defect(pred_orth)

    def corr_orthogonality(self, corr_orth):
        return _half_gram_defect(corr_orth)

    def pred_cov(self, A, sPp_tril, sPc_tril, sQ_tril, pred_orth):
        """sPp pred_orth = [A sPc, sQ] (time update, array form)."""
        rhs = np.hstack((A @ tril_mat(sPc_tril), tril_mat(sQ_tril)))
        return tril_mat(sPp_tril) @ pred_orth - rhs

    def corr_cov(self, C, sR_tril, sRp_tril, sPp_tril, sPc_tril, Kn,
                 corr_orth):
        """[[sRp,0],[Kn,sPc]] corr_orth = [[sR, C sPp],[0, sPp]]."""
        nx, ny = self.nx, self.ny
        sPp = tril_mat(sPp_tril)
        post = np.vstack((
            np.hstack((tril_mat(sRp_tril), np.zeros((ny, nx), dtype=int))),
            np.hstack((Kn, tril_mat(sPc_tril)))))
        pre = np.vstack((
            np.hstack((tril_mat(sR_tril), C @ sPp)),
            np.hstack((np.zeros((nx, ny), dtype=int), sPp))))
        return post @ corr_orth - pre

    def kalman_gain(self, Ln, Kn, A):
        return Ln - A @ Kn

    @property
    def generate_assignments(self):
        own = {'ntx': len(self.variables['sPp_tril'])}
        return {**own, **getattr(super(), 'generate_assignments', {})}


class ZOHDynamicsModel(InnovationDTModel):
    """Adds the zero-order-hold discretisation constraint (symfem.py:176-202)."""

    expm_order = 3

    def __init__(self, nx, nu, ny):
        super().__init__(nx, nu, ny)
        v = self.variables
        v['dt'] = 'dt'
        v['Ac'] = _matrix('Ac', nx, nx)
        v['Bc'] = _matrix('Bc', nx, nu)
        self.decision.update(('Ac', 'Bc'))
        self.add_constraint('discretize_AB')

    def discretize_AB(self, A, B, Ac, Bc, dt):
        """[A - expm_K(Ac dt), B - sum_{k=1..K} dt^k/k! Ac^(k-1) Bc]."""
        a_resid = A - expm_taylor(Ac * dt, self.expm_order)
        term = total = dt * Bc
        for k in range(2, self.expm_order + 1):
            term = (dt / k) * Ac @ term
            total = total + term
        return np.hstack((a_resid, B - total))


class DiscretizedNoiseModel(MaximumLikelihoodDTModel):
    """Adds the process-noise discretisation constraint (symfem.py:205-248)."""

    noise_disc_order = 1

    def __init__(self, nx, nu, ny):
        super().__init__(nx, nu, ny)
        v = self.variables
        v['sQc_tril'] = _tril('sQc', nx)
        if 'Ac' not in v:
            v['Ac'] = _matrix('Ac', nx, nx)
        if 'dt' not in v:
            v['dt'] = 'dt'
        self.decision.update(('Ac', 'sQc_tril'))
        self.add_constraint('discretize_Q')

    def discretize_Q(self, Ac, sQ_tril, sQc_tril, dt):
        """tril of sum_{j,k<=order} dt^(j+k+1)/((j+k+1) j! k!) Ac^j Qc Ac'^k - Q."""
        sQ, sQc = tril_mat(sQ_tril), tril_mat(sQc_tril)
        Qc = sQc @ sQc.T
        powers = [np.linalg.matrix_power(Ac, j)
                  for j in range(self.noise_disc_order + 1)]
        total = np.zeros_like(sQ)
        for j, Aj in enumerate(powers):
            for k, Ak in enumerate(powers):
                if k < j:
                    continue
                n = j + k + 1
                scal = dt ** n / n / float(math.factorial(j)) \
                    / float(math.factorial(k))
                term = scal * Aj @ Qc @ Ak.T
                total = total + (term if j == k else term + term.T)
        resid = total - sQ @ sQ.T
        return [resid[ij] for ij in tril_ind(self.nx)]


class TrapezoidalCTModel(symoptim.Model):
    """EXTENSION -- no reference counterpart at the reference's HEAD.

    Continuous-time innovation-form predictor ``dx/dt = Ac x + Bc u + Lc en``
    collocated with the trapezoidal rule (BASELINE.json names trapezoidal
    defects; the reference only has the discrete-time defect of
    symfem.py:50-53 and the parameter-only ZOH constraint of :193-202):

        x[k+1] - x[k] - dt/2 (f[k] + f[k+1]) = 0,
        f[k] = Ac x[k] + Bc u[k] + Lc en[k]

    a two-sample stencil in ``x``, ``en`` and ``u``.  ``innovation`` and
    ``loglikelihood`` are the reference's (symfem.py:55-65).
    """

    def __init__(self, nx, nu, ny):
        super().__init__()
        self.nx, self.nu, self.ny = nx, nu, ny
        v = self.variables
        for name, n in (('x', nx), ('en', ny), ('xnext', nx), ('xprev', nx),
                        ('enprev', ny), ('ennext', ny), ('ybias', ny)):
            v[name] = _vector(name, n)
        v['Ac'] = _matrix('Ac', nx, nx)
        v['Bc'] = _matrix('Bc', nx, nu)
        v['C'] = _matrix('C', ny, nx)
        v['D'] = _matrix('D', ny, nu)
        v['Lc'] = _matrix('Lc', nx, ny)
        v['sRp_tril'] = _tril('sRp', ny)
        self.decision.update(name for name in v if name != 'self')
        for name, n in (('u', nu), ('y', ny), ('uprev', nu), ('unext', nu)):
            v[name] = _vector(name, n)
        v['dt'] = 'dt'
        self.add_constraint('trapezoid')
        self.add_constraint('innovation')
        self.add_objective('loglikelihood')

    def trapezoid(self, xnext, xprev, unext, uprev, ennext, enprev, Ac, Bc,
                  Lc, dt):
        """Trapezoidal collocation defect of the continuous-time predictor."""
        fprev = Ac @ xprev + Bc @ uprev + Lc @ enprev
        fnext = Ac @ xnext + Bc @ unext + Lc @ ennext
        return xnext - xprev - 0.5 * dt * (fprev + fnext)

    innovation = InnovationDTModel.innovation
    loglikelihood = InnovationDTModel.loglikelihood

    @property
    def generate_assignments(self):
        return {'nx': self.nx, 'nu': self.nu, 'ny': self.ny,
                'nty': len(self.variables['sRp_tril'])}
