"""The trapezoidal-collocation extension (no reference counterpart): host
logic against its independent oracle statement and finite differences."""

import numpy as np
import scipy.sparse as sp

from colloc_fem_code_b200 import families
from oracle import ref_models


def _pair(N=6, dims=(2, 1, 2)):
    nx, nu, ny = dims
    rng = np.random.default_rng(21)
    y = rng.normal(size=(N, ny))
    u = rng.normal(size=(N, nu))
    return (families.make_problem('trapezoid', y, u, nx, dt=0.1),
            ref_models.make_problem('trapezoid', y, u, nx, dt=0.1), rng)


def test_layout_and_indices_match_oracle():
    p, o, _ = _pair()
    assert (p.ndec, p.ncons, p.nnzjac, p.nnzhess) == \
        (o.ndec, o.ncons, o.nnzjac, o.nnzhess)
    for a, b in zip(p.constr_jac_ind() + p.lag_hess_ind(),
                    o.constr_jac_ind() + o.lag_hess_ind()):
        np.testing.assert_array_equal(a, b)
    st = p.structure
    halo = {v['name']: v['hshift'] for v in st.vars if v['per_sample']}
    assert halo == {'x': 1, 'en': 1}          # two-sample stencil in both
    assert [(d['name'], d['hshift']) for d in st.data] == [('y', 0), ('u', 1)]


def test_oracle_derivatives_by_finite_differences():
    p, o, rng = _pair(N=4)
    dvec = rng.normal(size=o.ndec)
    o.variables(dvec)['sRp_tril'][[0, 2]] = [1.2, 0.7]
    lam = rng.normal(size=o.ncons)
    jr, jc = o.constr_jac_ind()
    hr, hc = o.lag_hess_ind()
    J = sp.coo_matrix((o.constr_jac_val(dvec), (jr, jc)),
                      shape=(o.ncons, o.ndec)).toarray()
    H = sp.coo_matrix((o.lag_hess_val(dvec, 0.8, lam), (hr, hc)),
                      shape=(o.ndec, o.ndec)).toarray()
    H = H + np.tril(H, -1).T

    def lag_grad(d):
        jv = sp.coo_matrix((o.constr_jac_val(d), (jr, jc)),
                           shape=(o.ncons, o.ndec))
        return 0.8 * o.obj_grad(d) + jv.T @ lam
    h = 1e-6
    for i in range(o.ndec):
        e = np.zeros(o.ndec)
        e[i] = h
        np.testing.assert_allclose(
            J[:, i], (o.constr(dvec + e) - o.constr(dvec - e)) / (2 * h),
            atol=1e-7)
        np.testing.assert_allclose(
            H[:, i], (lag_grad(dvec + e) - lag_grad(dvec - e)) / (2 * h),
            atol=1e-6)


def test_defect_vanishes_on_an_exact_trapezoidal_solution():
    """x[k+1] = (I - dt/2 Ac)^-1 ((I + dt/2 Ac) x[k] + dt/2 Bc (u[k]+u[k+1]))."""
    nx, nu, ny, N, dt = 2, 1, 2, 50, 0.05
    rng = np.random.default_rng(2)
    Ac = np.array([[-1.0, 0.5], [-0.3, -2.0]])
    Bc = rng.normal(size=(nx, nu))
    u = rng.normal(size=(N, nu))
    x = np.zeros((N, nx))
    M = np.linalg.inv(np.eye(nx) - dt / 2 * Ac)
    for k in range(N - 1):
        x[k + 1] = M @ ((np.eye(nx) + dt / 2 * Ac) @ x[k]
                        + dt / 2 * Bc @ (u[k] + u[k + 1]))
    o = ref_models.make_problem('trapezoid', x.copy(), u, nx, dt=dt)
    dvec = np.zeros(o.ndec)
    var = o.variables(dvec)
    var['Ac'][...] = Ac
    var['Bc'][...] = Bc
    var['C'][...] = np.eye(nx)
    var['x'][...] = x
    var['sRp_tril'][[0, 2]] = 1.0
    np.testing.assert_allclose(o.constr(dvec), 0, atol=1e-13)
