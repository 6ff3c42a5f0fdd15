"""Test helpers: an oracle-backed evaluator (CPU) for the NLP drivers and a
small identification problem in the style of attas_sp_innov.py."""

import numpy as np

from colloc_fem_code_b200 import nlp, synthetic


class OracleEvaluator(nlp.Evaluator):
    """Callbacks served by the CPU oracle (tests only)."""

    def __init__(self, oracle_problem):
        self.o = oracle_problem
        self.n, self.m = oracle_problem.ndec, oracle_problem.ncons
        self.seconds = 0.0
        self.calls = 0

    def jac_structure(self):
        return self.o.constr_jac_ind()

    def hess_structure(self):
        return self.o.lag_hess_ind()

    def time_structure(self):
        o = self.o
        dec = {n: (s.offset, s.shape) for n, s in o.decision.items()}
        con = {n: (r.spec.offset, r.spec.shape)
               for n, r in o.constraints.items()}
        return nlp.time_structure_of(dec, con, ('x', 'en'),
                                     {'dynamics': 1, 'innovation': 0})

    def eval_fg(self, x):
        self.calls += 1
        return self.o.obj(x), self.o.constr(x)

    def eval_all(self, x, sigma, lam):
        self.calls += 1
        return (self.o.obj(x), self.o.obj_grad(x), self.o.constr(x),
                self.o.constr_jac_val(x), self.o.lag_hess_val(x, sigma, lam))

    def ipopt_eval(self, which, x, new_x, out, sigma=None, lam=None):
        x = np.array(x)
        if which == 1:
            out[0] = self.o.obj(x)
        elif which == 2:
            out[:] = self.o.obj_grad(x)
        elif which == 4:
            out[:] = self.o.constr(x)
        elif which == 8:
            out[:] = self.o.constr_jac_val(x)
        else:
            out[:] = self.o.lag_hess_val(x, sigma, np.array(lam))


def attas_like_experiment(seed, N, nx=2, nu=1, sw=0.05, sv=0.1):
    """Synthetic stand-in for the ATTAS short-period data: every state is
    measured (C = I, D = 0, as the scripts impose through bounds)."""
    rng = np.random.default_rng(seed)
    A, B, _, _ = synthetic.random_stable_system(rng, nx, nu, nx, rho=0.9)
    C, D = np.eye(nx), np.zeros((nx, nu))
    u, y, x = synthetic.simulate(rng, N, A, B, C, D, std_w=sw, std_v=sv)
    return {'A': A, 'B': B, 'C': C, 'D': D, 'u': u, 'y': y, 'x': x}


def innovation_setup(problem, exp, tril_diag, feasible=True):
    """Initial guess, bounds, scaling of /root/reference/attas_sp_innov.py:
    88-135: previous-sample predictor states, A = I, B = 0, C = I and D = 0
    fixed through equal bounds, sRp diagonal bounded below."""
    y, u = exp['y'], exp['u']
    nx = exp['A'].shape[0]
    ny = y.shape[1]
    dec0 = np.zeros(problem.ndec)
    var0 = problem.variables(dec0)
    if feasible:
        # predictor-consistent start (the "predict" step sketched in
        # /root/reference/mc_blackbox_cfem.py:53-74): a rough parameter
        # guess, states and innovations from running the predictor
        rng = np.random.default_rng(1)
        A0 = exp['A'] * (1 + 0.2 * rng.normal(size=(nx, nx)))
        B0 = exp['B'] * (1 + 0.2 * rng.normal(size=exp['B'].shape))
        from colloc_fem_code_b200 import fit
        guess = fit.predictor_guess(y, u, A0, B0, exp['C'], exp['D'],
                                    np.zeros((nx, ny)))
        for k, v in guess.items():
            var0[k][...] = v
    else:
        x0 = np.vstack((np.zeros(nx), y[:-1, :nx]))
        Rp0 = np.cov(y - x0 @ exp['C'].T, rowvar=0).reshape(ny, ny)
        sRp0 = np.linalg.cholesky(Rp0)
        en0 = np.linalg.solve(sRp0, (y - x0 @ exp['C'].T).T).T
        var0['A'][...] = np.eye(nx)
        var0['C'][...] = exp['C']
        var0['x'][...] = x0
        var0['en'][...] = en0
        var0['sRp_tril'][...] = sRp0[np.tril_indices(ny)]
    dec_bounds = np.repeat([[-np.inf], [np.inf]], problem.ndec, axis=-1)
    var_L = problem.variables(dec_bounds[0])
    var_U = problem.variables(dec_bounds[1])
    var_L['sRp_tril'][tril_diag(ny)] = 1e-7
    var_L['C'][...] = exp['C']
    var_U['C'][...] = exp['C']
    var_L['D'][...] = 0
    var_U['D'][...] = 0
    constr_bounds = np.zeros((2, problem.ncons))
    constr_scale = np.ones(problem.ncons)
    problem.unpack_constraints(constr_scale)['innovation'][...] = 1e2
    dec_scale = np.ones(problem.ndec)
    var_scale = problem.variables(dec_scale)
    var_scale['sRp_tril'][...] = 1e2
    var_scale['Ln'][...] = 1e2
    return dec0, dec_bounds, constr_bounds, (-1.0, dec_scale, constr_scale)
