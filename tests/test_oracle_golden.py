"""The CPU oracle (oracle/ref_models.py restatement on oracle/engine.py)
against the committed golden fixtures, which were generated from the
reference's own unchanged symfem.py / fem.py (oracle/make_golden.py)."""

import numpy as np

from oracle import ref_models

RTOL = 1e-13


def _problem(g):
    nx, nu, ny = g['dims']
    return ref_models.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])


def test_oracle_reproduces_golden(golden):
    g = golden
    p = _problem(g)
    assert (p.ndec, p.ncons) == (g['ndec'], g['ncons'])
    jr, jc = p.constr_jac_ind()
    hr, hc = p.lag_hess_ind()
    np.testing.assert_array_equal(jr, g['jac_row'])
    np.testing.assert_array_equal(jc, g['jac_col'])
    np.testing.assert_array_equal(hr, g['hess_row'])
    np.testing.assert_array_equal(hc, g['hess_col'])
    dvec = g['dvec']
    np.testing.assert_allclose(p.obj(dvec), g['f'], rtol=RTOL)
    np.testing.assert_allclose(p.obj_grad(dvec), g['grad'], rtol=RTOL,
                               atol=1e-15)
    np.testing.assert_allclose(p.constr(dvec), g['g'], rtol=RTOL, atol=1e-14)
    np.testing.assert_allclose(p.constr_jac_val(dvec), g['jac_val'],
                               rtol=RTOL, atol=1e-15)
    np.testing.assert_allclose(
        p.lag_hess_val(dvec, g['obj_factor'], g['lam']), g['hess_val'],
        rtol=RTOL, atol=1e-15)


def test_known_answer_noise_free():
    """Noise-free LTI data, true states, zero innovations => all defects vanish
    and the objective is -N log det sRp (symfem.py:50-65)."""
    rng = np.random.default_rng(5)
    nx, nu, ny, N = 3, 2, 2, 40
    A = np.diag(rng.uniform(-0.8, 0.8, nx))
    B = rng.normal(size=(nx, nu))
    C = rng.normal(size=(ny, nx))
    D = rng.normal(size=(ny, nu))
    u = rng.normal(size=(N, nu))
    x = np.zeros((N, nx))
    for k in range(N - 1):
        x[k + 1] = A @ x[k] + B @ u[k]
    y = x @ C.T + u @ D.T
    p = ref_models.make_problem('innovation', y, u, nx)
    dvec = np.zeros(p.ndec)
    var = p.variables(dvec)
    var['A'][:] = A
    var['B'][:] = B
    var['C'][:] = C
    var['D'][:] = D
    var['x'][:] = x
    sRp = np.array([1.5, 0.3, 0.7])
    var['sRp_tril'][:] = sRp
    np.testing.assert_allclose(p.constr(dvec), 0, atol=1e-12)
    np.testing.assert_allclose(p.obj(dvec), -N * np.log(1.5 * 0.7),
                               rtol=1e-14)


def test_derivatives_by_finite_differences():
    """Jacobian and Lagrangian Hessian of the oracle against central
    differences of its own values / gradients."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    nx, nu, ny, N = 2, 1, 2, 4
    y = rng.normal(size=(N, ny))
    u = rng.normal(size=(N, nu))
    p = ref_models.make_problem('ndisc_zoh', y, u, nx, dt=0.1)
    dvec = rng.normal(size=p.ndec)
    for name, s in p.decision.items():
        if name.endswith('_tril'):
            s.unpack_from(dvec)[ref_models.tril_diag(
                int(round((np.sqrt(8 * s.size + 1) - 1) / 2)))] = \
                rng.uniform(0.5, 2, size=int(round(
                    (np.sqrt(8 * s.size + 1) - 1) / 2)))
    lam = rng.normal(size=p.ncons)
    sigma = 0.7
    jr, jc = p.constr_jac_ind()
    J = sp.coo_matrix((p.constr_jac_val(dvec), (jr, jc)),
                      shape=(p.ncons, p.ndec)).toarray()
    hr, hc = p.lag_hess_ind()
    H = sp.coo_matrix((p.lag_hess_val(dvec, sigma, lam), (hr, hc)),
                      shape=(p.ndec, p.ndec)).toarray()
    H = H + np.tril(H, -1).T
    h = 1e-6
    Jfd = np.empty_like(J)
    Hfd = np.empty_like(H)

    def lag_grad(d):
        jv = sp.coo_matrix((p.constr_jac_val(d), (jr, jc)),
                           shape=(p.ncons, p.ndec))
        return sigma * p.obj_grad(d) + jv.T @ lam

    for i in range(p.ndec):
        e = np.zeros(p.ndec)
        e[i] = h
        Jfd[:, i] = (p.constr(dvec + e) - p.constr(dvec - e)) / (2 * h)
        Hfd[:, i] = (lag_grad(dvec + e) - lag_grad(dvec - e)) / (2 * h)
    np.testing.assert_allclose(J, Jfd, atol=1e-7)
    np.testing.assert_allclose(H, Hfd, atol=1e-6)
