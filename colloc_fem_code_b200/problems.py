"""Filter-error problem families (host-side mirror of /root/reference/fem.py).

Same class names, constructor ``(model, y, u)``, decision / constraint
registration order and cooperative multiple inheritance as the reference, so
``class Problem(fem.MaximumLikelihoodDTProblem, fem.ZOHDynamicsProblem)``
(/root/reference/attas_sp_ml_zoh.py:53-54) composes exactly as it does there.
The flat layouts that result are listed in SURVEY.md appendix B and pinned by
``tests/golden``.  Evaluation of everything registered here happens in the
fused CUDA kernels (``optim.Problem.backend``).
"""

import numpy as np

from . import optim


def _ntril(n):
    return n * (n + 1) // 2


class InnovationDTProblem(optim.Problem):
    """Predictor states + normalised innovations as decisions (fem.py:9-62)."""

    def __init__(self, model, y, u):
        super().__init__()
        self.model = model
        self.y = np.asarray(y)
        self.u = np.asarray(u)
        self.uprev = self.u[:-1]        # view: shares memory with u
        self.N = N = len(self.y)
        nx, nu, ny = model.nx, model.nu, model.ny
        if self.y.ndim != 2 or self.y.shape[1] != ny:
            raise AssertionError(f'y must be (N, {ny})')
        if self.u.shape != (N, nu):
            raise AssertionError(f'u must be ({N}, {nu})')
        if N <= 1:
            raise AssertionError('need at least two samples')

        # parameters first, then the per-sample slabs (fem.py:36-44)
        for name, shape in (('ybias', ny), ('sRp_tril', model.nty),
                            ('A', (nx, nx)), ('B', (nx, nu)), ('C', (ny, nx)),
                            ('D', (ny, nu)), ('Ln', (nx, ny))):
            self.add_decision(name, shape)
        x = self.add_decision('x', (N, nx))
        en = self.add_decision('en', (N, ny))

        # one-sample-shifted views of the slabs (fem.py:47-52)
        shifted = {'xprev': ((N - 1, nx), x.offset),
                   'enprev': ((N - 1, ny), en.offset),
                   'xnext': ((N - 1, nx), x.offset + nx)}
        for name, (shape, offset) in shifted.items():
            self.add_dependent_variable(name, optim.Decision(shape, offset))

        self.add_objective(model.loglikelihood, N)
        self.add_constraint(model.dynamics, (N - 1, nx))
        self.add_constraint(model.innovation, (N, ny))

    def variables(self, dvec):
        """Decision views plus the data the model functions read."""
        data = {'y': self.y, 'u': self.u, 'uprev': self.uprev}
        return {**data, **super().variables(dvec)}


class BalancedDTProblem(InnovationDTProblem):
    """Balanced-realisation constraints (fem.py:66-84)."""

    def __init__(self, model, y, u):
        super().__init__(model, y, u)
        nx, nu, ny = model.nx, model.nu, model.ny
        self.add_decision('sW_diag', nx)
        self.add_decision('ctrl_orth', (nx, nx + nu))
        self.add_decision('obs_orth', (nx, nx + ny))
        self.add_constraint(model.ctrl_gram, (nx, nx + nu))
        self.add_constraint(model.obs_gram, (nx, nx + ny))
        self.add_constraint(model.ctrl_orthogonality, _ntril(nx))
        self.add_constraint(model.obs_orthogonality, _ntril(nx))


class MaximumLikelihoodDTProblem(InnovationDTProblem):
    """Square-root Riccati constraints (fem.py:87-109)."""

    def __init__(self, model, y, u):
        super().__init__(model, y, u)
        nx, ny = model.nx, model.ny
        nxy = nx + ny
        for name in ('sPp_tril', 'sPc_tril', 'sQ_tril'):
            self.add_decision(name, model.ntx)
        self.add_decision('sR_tril', model.nty)
        self.add_decision('Kn', (nx, ny))
        self.add_decision('pred_orth', (nx, 2 * nx))
        self.add_decision('corr_orth', (nxy, nxy))
        self.add_constraint(model.pred_orthogonality, model.ntx)
        self.add_constraint(model.corr_orthogonality, _ntril(nxy))
        self.add_constraint(model.pred_cov, (nx, 2 * nx))
        self.add_constraint(model.corr_cov, (nxy, nxy))
        self.add_constraint(model.kalman_gain, (nx, ny))


def _with_sample_time(problem, variables):
    """``dt`` is an auxiliary scalar read off the model (fem.py:126-128)."""
    variables['dt'] = problem.model.dt
    return variables


class ZOHDynamicsProblem(InnovationDTProblem):
    """Zero-order-hold discretisation constraint (fem.py:112-128)."""

    def __init__(self, model, y, u):
        super().__init__(model, y, u)
        nx, nu = model.nx, model.nu
        self.add_decision('Ac', (nx, nx))
        self.add_decision('Bc', (nx, nu))
        self.add_constraint(model.discretize_AB, (nx, nx + nu))

    def variables(self, dvec):
        return _with_sample_time(self, super().variables(dvec))


class DiscretizedNoiseProblem(MaximumLikelihoodDTProblem):
    """Process-noise discretisation constraint (fem.py:131-149)."""

    def __init__(self, model, y, u):
        super().__init__(model, y, u)
        nx = model.nx
        self.add_decision('sQc_tril', model.ntx)
        if 'Ac' not in self.decision:
            self.add_decision('Ac', (nx, nx))
        self.add_constraint(model.discretize_Q, _ntril(nx))

    def variables(self, dvec):
        return _with_sample_time(self, super().variables(dvec))


class TrapezoidalCTProblem(optim.Problem):
    """EXTENSION -- no reference counterpart (see models.TrapezoidalCTModel).
    Layout in the manner of fem.py:36-57: parameters, then the ``x`` and
    ``en`` slabs; ``xprev/xnext`` and ``enprev/ennext`` are the two one-sample
    shifted views of the slabs, ``uprev/unext`` those of the inputs."""

    def __init__(self, model, y, u):
        super().__init__()
        self.model = model
        self.y = np.asarray(y)
        self.u = np.asarray(u)
        self.uprev, self.unext = self.u[:-1], self.u[1:]
        self.N = N = len(self.y)
        nx, nu, ny = model.nx, model.nu, model.ny
        if self.y.shape != (N, ny) or self.u.shape != (N, nu) or N <= 1:
            raise AssertionError('y must be (N, ny), u (N, nu), N > 1')
        for name, shape in (('ybias', ny), ('sRp_tril', model.nty),
                            ('Ac', (nx, nx)), ('Bc', (nx, nu)),
                            ('C', (ny, nx)), ('D', (ny, nu)),
                            ('Lc', (nx, ny))):
            self.add_decision(name, shape)
        x = self.add_decision('x', (N, nx))
        en = self.add_decision('en', (N, ny))
        for name, shape, offset in (
                ('xprev', (N - 1, nx), x.offset),
                ('xnext', (N - 1, nx), x.offset + nx),
                ('enprev', (N - 1, ny), en.offset),
                ('ennext', (N - 1, ny), en.offset + ny)):
            self.add_dependent_variable(name, optim.Decision(shape, offset))
        self.add_objective(model.loglikelihood, N)
        self.add_constraint(model.trapezoid, (N - 1, nx))
        self.add_constraint(model.innovation, (N, ny))

    def variables(self, dvec):
        data = {'y': self.y, 'u': self.u, 'uprev': self.uprev,
                'unext': self.unext, 'dt': self.model.dt}
        return {**data, **super().variables(dvec)}
