#!/usr/bin/env python3
"""The flow of the reference's ``attas_sp_ml.py`` on the B200 path.

Same steps as /root/reference/attas_sp_ml.py:79-184 -- build the symbolic
model, compile it, build the problem, initial guess, bounds, scaling, solve
through ``problem.ipopt(...)``, unpack -- with two differences forced by the
environment: the flight-test file ``data/fAttasElv1.mat`` is not shipped, so
the data are a synthetic short-period-like record (every state measured), and
the start is the feasible steady-state-filter point of ``fit.kalman_guess``
instead of the previous-sample predictor.

    python examples/attas_sp_ml_synthetic.py [N]
"""

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import colloc_fem_code_b200.compat as compat  # noqa: E402

compat.install(mirrors=True)        # `import fem, symfem` -> the GPU-backed mirrors

import fem  # noqa: E402
import symfem  # noqa: E402
from colloc_fem_code_b200 import fit, synthetic  # noqa: E402


def load_data(N, seed=0):
    """Synthetic stand-in for attas_sp_ml.py:21-46 (2 states, 1 input, both
    states measured, seed-0 noise)."""
    rng = np.random.default_rng(seed)
    A, B, _, _ = synthetic.random_stable_system(rng, 2, 1, 2, rho=0.9)
    C, D = np.eye(2), np.zeros((2, 1))
    u, y, x = synthetic.simulate(rng, N, A, B, C, D, std_w=0.1, std_v=0.1)
    return u, y, (A, B, C, D)


if __name__ == '__main__':
    nx, nu, ny = 2, 1, 2
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    u, y, (A, B, C, D) = load_data(N)

    t0 = time.perf_counter()
    symmodel = symfem.MaximumLikelihoodDTModel(nx=nx, nu=nu, ny=ny)
    model = symmodel.compile_class()()
    problem = fem.MaximumLikelihoodDTProblem(model, y, u)
    print(f'model + problem: {time.perf_counter() - t0:.2f} s  '
          f'(ndec {problem.ndec}, ncons {problem.ncons})')

    # initial guess: equation-error estimate of A, B (attas_sp_ml.py:49-75)
    Z = np.hstack((y[:-1], u[:-1]))
    est = np.linalg.lstsq(Z, y[1:], rcond=None)[0].T
    A0, B0 = est[:, :nx], est[:, nx:]
    guess = fit.kalman_guess(y, u, A0, B0, C, D, 0.1 * np.eye(nx),
                             0.1 * np.eye(ny))
    dec0 = np.zeros(problem.ndec)
    var0 = problem.variables(dec0)
    for name, value in guess.items():
        var0[name][:] = value

    # bounds (attas_sp_ml.py:113-131)
    dec_bounds = np.repeat([[-np.inf], [np.inf]], problem.ndec, axis=-1)
    dec_L, dec_U = dec_bounds
    var_L = problem.variables(dec_L)
    var_U = problem.variables(dec_U)
    var_L['C'][:] = np.eye(2)
    var_U['C'][:] = np.eye(2)
    var_L['D'][:] = np.zeros((2, 1))
    var_U['D'][:] = np.zeros((2, 1))
    var_L['sRp_tril'][symfem.tril_diag(2)] = 1e-6
    var_L['sPp_tril'][symfem.tril_diag(2)] = 0
    var_L['sPc_tril'][symfem.tril_diag(2)] = 0
    var_L['sQ_tril'][symfem.tril_diag(2)] = 0
    var_L['sR_tril'][symfem.tril_diag(2)] = 1e-6
    var_L['sR_tril'][~symfem.tril_diag(2)] = 0
    var_U['sR_tril'][~symfem.tril_diag(2)] = 0
    constr_bounds = np.zeros((2, problem.ncons))

    # scaling (attas_sp_ml.py:134-151)
    obj_scale = -1.0
    constr_scale = np.ones(problem.ncons)
    var_constr_scale = problem.unpack_constraints(constr_scale)
    for name in ('innovation', 'pred_cov', 'corr_cov', 'kalman_gain'):
        var_constr_scale[name][:] = 100
    dec_scale = np.ones(problem.ndec)
    var_scale = problem.variables(dec_scale)
    for name in ('Ln', 'sRp_tril', 'sPp_tril', 'sPc_tril', 'sQ_tril',
                 'sR_tril', 'Kn'):
        var_scale[name][:] = 1e2

    with problem.ipopt(dec_bounds, constr_bounds) as nlp:
        nlp.add_str_option('linear_solver', 'ma57')
        nlp.add_num_option('ma57_pre_alloc', 20.0)
        nlp.add_num_option('tol', 1e-8)
        nlp.add_int_option('max_iter', 1000)
        nlp.set_scaling(obj_scale, dec_scale, constr_scale)
        decopt, info = nlp.solve(dec0)

    opt = problem.variables(decopt)
    sRp = symfem.tril_mat(opt['sRp_tril'])
    sQ = symfem.tril_mat(opt['sQ_tril'])
    sR = symfem.tril_mat(opt['sR_tril'])
    print('solver', info['solver'], 'status', info['status'], 'iterations',
          info['iterations'])
    print(f"time: total {info['seconds_total']:.2f} s, callbacks (GPU path) "
          f"{info['seconds_callbacks']:.3f} s, KKT (host) "
          f"{info['seconds_kkt']:.2f} s")
    print('A estimate\n', opt['A'], '\nA true\n', A)
    print('B estimate', opt['B'].ravel(), 'true', B.ravel())
    print('sQ diag', np.diag(sQ), '(true 0.1)  sR diag', np.diag(sR),
          '(true 0.1)  sRp diag', np.diag(sRp))
