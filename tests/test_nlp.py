"""NLP drivers on the CPU: the built-in interior-point method, its structured
KKT solver, the request-yielding iteration used for lock-step batches, and
the ctypes binding of IPOPT's C interface (against a fake libipopt).  The
callbacks come from the CPU oracle here; the GPU-backed versions are in
test_gpu_nlp.py."""

import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from colloc_fem_code_b200 import families, fit, models, nlp
from oracle import ref_models

from nlp_helpers import OracleEvaluator, attas_like_experiment

HERE = os.path.dirname(os.path.abspath(__file__))


class HS71(nlp.Evaluator):
    """Hock-Schittkowski 71 with a slack for the inequality."""
    n, m = 5, 2

    def jac_structure(self):
        return (np.array([0, 0, 0, 0, 1, 1, 1, 1, 1]),
                np.array([0, 1, 2, 3, 0, 1, 2, 3, 4]))

    def hess_structure(self):
        return np.tril_indices(4)

    def eval_fg(self, x):
        return (x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2],
                np.array([x @ x - x[4] ** 2 - 40,
                          x[0] * x[1] * x[2] * x[3] - x[4] - 25]))

    def eval_all(self, x, sigma, lam):
        f, g = self.eval_fg(x)
        grad = np.array([x[3] * (2 * x[0] + x[1] + x[2]), x[0] * x[3],
                         x[0] * x[3] + 1, x[0] * (x[0] + x[1] + x[2]), 0.0])
        jac = np.array([2 * x[0], 2 * x[1], 2 * x[2], 2 * x[3],
                        x[1] * x[2] * x[3], x[0] * x[2] * x[3],
                        x[0] * x[1] * x[3], x[0] * x[1] * x[2], -1.0])
        H = np.zeros((4, 4))
        H[0, 0], H[1, 0], H[2, 0] = 2 * x[3], x[3], x[3]
        H[3, 0], H[3, 1], H[3, 2] = 2 * x[0] + x[1] + x[2], x[0], x[0]
        H *= sigma
        H += lam[0] * 2 * np.eye(4)
        P = np.zeros((4, 4))
        P[1, 0], P[2, 0], P[3, 0] = x[2] * x[3], x[1] * x[3], x[1] * x[2]
        P[2, 1], P[3, 1], P[3, 2] = x[0] * x[3], x[0] * x[2], x[0] * x[1]
        H += lam[1] * P
        return f, grad, g, jac, H[np.tril_indices(4)]


def test_builtin_ipm_solves_hs71():
    db = np.array([[1, 1, 1, 1, 0], [5, 5, 5, 5, np.inf]], float)
    s = nlp.InteriorPointSolver(HS71(), db, np.zeros((2, 2)))
    s.add_num_option('tol', 1e-9)
    x, info = s.solve(np.array([1, 5, 5, 1, 0.0]))
    assert info['status'] == 'solved'
    np.testing.assert_allclose(info['obj'], 17.0140173, rtol=1e-7)
    np.testing.assert_allclose(x[:4], [1, 4.7429994, 3.8211503, 1.3794082],
                               rtol=1e-6)


def _ml_case(N=150, nx=2):
    exp = attas_like_experiment(7, N, sw=0.1)
    o = ref_models.make_problem('ml', exp['y'], exp['u'], nx)
    p = families.make_problem('ml', exp['y'], exp['u'], nx)
    rng = np.random.default_rng(1)
    A0 = exp['A'] * (1 + 0.1 * rng.normal(size=(nx, nx)))
    B0 = exp['B'] * (1 + 0.1 * rng.normal(size=exp['B'].shape))
    guess = fit.kalman_guess(exp['y'], exp['u'], A0, B0, exp['C'], exp['D'],
                             0.1 * np.eye(nx), 0.2 * np.eye(nx))
    dec0 = fit.start_point(p, guess)
    bounds = fit.ml_setup(p, fix={'C': exp['C'], 'D': exp['D']})
    return exp, o, p, dec0, bounds


def test_kalman_guess_is_feasible():
    exp, o, p, dec0, _ = _ml_case()
    assert np.max(np.abs(o.constr(dec0))) < 1e-12


def test_structured_kkt_inertia_matches_dense_eigenvalues():
    exp, o, p, dec0, _ = _ml_case(N=30)
    rng = np.random.default_rng(0)
    x = rng.normal(size=o.ndec)
    lam = rng.normal(size=o.ncons)
    jr, jc = o.constr_jac_ind()
    hr, hc = o.lag_hess_ind()
    n, m = o.ndec, o.ncons
    J = sp.csr_matrix((o.constr_jac_val(x), (jr, jc)), shape=(m, n))
    W = sp.coo_matrix((o.lag_hess_val(x, 1.0, lam), (hr, hc)), shape=(n, n))
    W = (W + sp.tril(W, -1).T).tocsr()
    kkt = nlp.BorderedBandKKT(*OracleEvaluator(o).time_structure())
    rhs = rng.normal(size=n + m)
    for dw in (0.0, 1.0, 1e3):
        K = sp.bmat([[W + dw * sp.identity(n), J.T], [J, None]]).toarray()
        sol, _, neg = kkt.solve(W, J, dw, 0.0, rhs, want_inertia=True)
        assert neg == int((np.linalg.eigvalsh(K) < 0).sum())
        np.testing.assert_allclose(K @ sol, rhs, atol=1e-8 * np.abs(rhs).max())


def test_ml_fit_converges_and_generator_equals_direct_solve():
    exp, o, p, dec0, (db, cb, scaling) = _ml_case()
    ev = OracleEvaluator(o)
    s = nlp.InteriorPointSolver(ev, db, cb)
    s.add_num_option('tol', 1e-8)
    s.set_scaling(*scaling)
    x, info = s.solve(dec0)
    assert info['status'] == 'solved', info['status']
    assert np.max(np.abs(o.constr(x))) < 1e-8
    # the same iteration driven from outside (what fit.BatchFitter does)
    s2 = nlp.InteriorPointSolver(ev, db, cb)
    s2.add_num_option('tol', 1e-8)
    s2.set_scaling(*scaling)
    steps = s2.solve_steps(dec0)
    req = next(steps)
    while True:
        res = ev.eval_all(*req[1:]) if req[0] == 'all' else ev.eval_fg(req[1])
        try:
            req = steps.send(res)
        except StopIteration as stop:
            x2, info2 = stop.value
            break
    np.testing.assert_array_equal(x, x2)
    assert info2['iterations'] == info['iterations']


@pytest.fixture(scope='module')
def fake_ipopt(tmp_path_factory):
    out = tmp_path_factory.mktemp('fake') / 'libipopt_fake.so'
    subprocess.run(['gcc', '-shared', '-fPIC', '-O1', '-o', str(out),
                    os.path.join(HERE, 'fake_ipopt.c')], check=True)
    return str(out)


def test_ipopt_binding_against_fake_library(fake_ipopt):
    exp, o, p, dec0, (db, cb, scaling) = _ml_case(N=12)
    ev = OracleEvaluator(o)
    s = nlp.IpoptSolver(ev, db, cb, libpath=fake_ipopt)
    s.add_str_option('linear_solver', 'ma57')
    s.add_num_option('tol', 1e-7)
    s.add_int_option('max_iter', 123)
    s.set_scaling(*scaling)
    x, info = s.solve(dec0)
    s.close()
    assert info['status'] == 0 and info['last_error'] is None
    lam = 0.5 + 0.01 * np.arange(o.ncons)
    jr, jc = o.constr_jac_ind()
    hr, hc = o.lag_hess_ind()
    np.testing.assert_allclose(info['obj'], o.obj(dec0), rtol=1e-14)
    np.testing.assert_allclose(info['g'], o.constr(dec0), atol=1e-14)
    np.testing.assert_allclose(info['mult_g'], lam)
    L, U = info['mult_x_L'], info['mult_x_U']
    np.testing.assert_allclose(L[0], o.constr_jac_val(dec0).sum(), rtol=1e-12)
    np.testing.assert_allclose(L[1], o.lag_hess_val(dec0, 0.75, lam).sum(),
                               rtol=1e-12)
    np.testing.assert_allclose(L[2], o.obj_grad(dec0).sum(), rtol=1e-12)
    np.testing.assert_allclose(
        L[3], (jr + 2.0 * jc).sum() + (3.0 * hr + 5.0 * hc).sum())
    assert (U[0], U[1], U[2], U[4]) == (1e-7, 123, -1.0, 1.0)
    np.testing.assert_allclose(U[3], scaling[1][0] + 10 * scaling[2][0])


def test_ipopt_binding_reports_evaluation_errors(fake_ipopt):
    exp, o, p, dec0, (db, cb, scaling) = _ml_case(N=12)

    class Broken(OracleEvaluator):
        def ipopt_eval(self, which, *args, **kw):
            if which == 8:
                raise RuntimeError('device lost')
            return super().ipopt_eval(which, *args, **kw)
    s = nlp.IpoptSolver(Broken(o), db, cb, libpath=fake_ipopt)
    x, info = s.solve(dec0)
    s.close()
    assert info['status'] == -13
    assert 'device lost' in str(info['last_error'])


def test_parallel_batch_fitter_does_not_hang_when_the_parent_fails():
    """No CUDA device here: cfem_create fails in the parent after the workers
    were forked.  The barrier must be aborted so that the workers exit and the
    error surfaces (it used to block for ever)."""
    import time
    from colloc_fem_code_b200 import backend, families, fit, synthetic
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('needs a box WITHOUT a CUDA device')
    except ImportError:
        pass
    nx, nu, ny, N = 2, 1, 2, 40
    problems = []
    for s in range(2):
        exp = synthetic.experiment(s, N, nx, nu, ny)
        problems.append(families.make_problem('innovation', exp['y'],
                                              exp['u'], nx))
    p = problems[0]
    pf = fit.ParallelBatchFitter(problems, workers=2)
    bounds = np.repeat([[-np.inf], [np.inf]], p.ndec, axis=-1)
    t0 = time.perf_counter()
    with pytest.raises(backend.CfemError):
        pf.fit([np.zeros(p.ndec)] * 2, bounds, np.zeros((2, p.ncons)),
               (-1.0, None, None), max_iter=3)
    assert time.perf_counter() - t0 < 60


def test_batch_needs_one_obj_scale():
    from colloc_fem_code_b200 import fit
    assert fit._common_obj_scale([-1.0, -1.0]) == -1.0
    with pytest.raises(ValueError):
        fit._common_obj_scale([-1.0, 1.0])


def test_fit_with_retries_only_retries_what_failed():
    """fit.fit_with_retries: every attempt is one batch of the problems that
    have not solved yet; the result keeps the attempt that solved it."""
    from colloc_fem_code_b200 import fit
    calls = []

    class FakeFitter:
        launches, seconds_gpu = 7, 0.5

        def __init__(self, sub):
            self.sub = sub

        def fit(self, dec0s, dec_bounds, constr_bounds, scaling, tol,
                max_iter, options):
            calls.append(([p['id'] for p in self.sub], dict(options or {}),
                          float(dec_bounds[0][0])))
            out = []
            for p in self.sub:
                ok = p['solves_at'] <= len(calls) - 1
                out.append((np.full(2, float(p['id'])),
                            {'status': 'solved' if ok else 'max_iter'}))
            return out

        def close(self):
            pass

    class P(dict):
        def variables(self, vec):
            return {'sQ_tril': vec[:1], 'other': vec[1:]}

    problems = [P(id=0, solves_at=0), P(id=1, solves_at=3),
                P(id=2, solves_at=1), P(id=3, solves_at=9)]
    bounds = np.array([[0.0, -1.0], [np.inf, np.inf]])
    res, report = fit.fit_with_retries(FakeFitter, problems, [None] * 4,
                                       bounds, None, None,
                                       attempts=fit.MC_ATTEMPTS[:4])
    assert [c[0] for c in calls] == [[0, 1, 2, 3], [1, 2, 3], [1, 3], [1, 3]]
    assert calls[1][1] == {'mu_init': 1e-2} and calls[3][2] == -np.inf
    assert [r[1]['attempt'] for r in res] == [0, 3, 1, 3]
    assert [fit.solved(r[1]) for r in res] == [True, True, True, False]
    assert [r['solved'] for r in report] == [1, 1, 0, 1]
    assert bounds[0][0] == 0.0              # the caller's bounds are untouched


def test_staged_start_is_feasible_for_the_ml_problem():
    """fit.staged_start: from a point of the BalancedDT problem (here: the
    predictor run of a balanced model) to a start of ML+Balanced that satisfies
    every constraint of that problem (oracle evaluation)."""
    from colloc_fem_code_b200 import fit, synthetic
    nx, nu, ny, N = 3, 2, 2, 120
    exp = synthetic.experiment(5, N, nx, nu, ny)
    y, u = exp['y'], exp['u']
    T, bal = fit.balanced_guess(exp['A'], exp['B'], exp['C'])
    g = fit.kalman_guess(y, u, bal['A'], bal['B'], bal['C'], exp['D'],
                         0.05 * np.eye(nx), 0.2 * np.eye(ny))
    g.update({k: bal[k] for k in ('sW_diag', 'ctrl_orth', 'obs_orth')})
    pb = families.make_problem('balanced', y, u, nx)
    pm = families.make_problem('ml_balanced', y, u, nx)
    bal_dec = fit.start_point(pb, g)
    ob = ref_models.make_problem('balanced', y, u, nx)
    assert np.max(np.abs(ob.constr(bal_dec))) < 1e-9
    dec0 = fit.staged_start(pm, pb, bal_dec)
    om = ref_models.make_problem('ml_balanced', y, u, nx)
    assert np.max(np.abs(om.constr(dec0))) < 1e-8
    db, cb, _ = fit.ml_setup(pm)
    assert np.all(dec0 >= db[0]) and np.all(dec0 <= db[1])
    dbb, cbb, scb = fit.balanced_setup(pb)
    assert dbb.shape == (2, pb.ndec) and np.all(bal_dec >= dbb[0])
