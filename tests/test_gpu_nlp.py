"""Solution-level parity (BASELINE north star: the same solution to 1e-8
relative on the parameters): the same NLP driver is run once with callbacks
from the CUDA path and once with callbacks from the CPU oracle.  IPOPT is not
installed in this image, so the driver is the built-in interior-point method
(colloc_fem_code_b200/nlp.py); with a libipopt present `problem.ipopt` binds it
instead."""

import numpy as np
import pytest

from colloc_fem_code_b200 import families, fit, nlp
from oracle import ref_models

from nlp_helpers import OracleEvaluator, attas_like_experiment

pytestmark = pytest.mark.gpu

PARAMS = ('A', 'B', 'Ln', 'ybias', 'sRp_tril', 'sQ_tril', 'sR_tril', 'Kn',
          'sPp_tril', 'sPc_tril')


def _case(seed, N, kind='ml', nx=2, sw=0.1):
    exp = attas_like_experiment(seed, N, sw=sw)
    p = families.make_problem(kind, exp['y'], exp['u'], nx)
    o = ref_models.make_problem(kind, exp['y'], exp['u'], nx)
    rng = np.random.default_rng(1)
    A0 = exp['A'] * (1 + 0.1 * rng.normal(size=(nx, nx)))
    B0 = exp['B'] * (1 + 0.1 * rng.normal(size=exp['B'].shape))
    guess = fit.kalman_guess(exp['y'], exp['u'], A0, B0, exp['C'], exp['D'],
                             0.1 * np.eye(nx), 0.2 * np.eye(nx))
    dec0 = fit.start_point(p, guess)
    setup = fit.ml_setup(p, fix={'C': exp['C'], 'D': exp['D']})
    return exp, p, o, dec0, setup


def test_same_solution_as_oracle_driven_solve():
    exp, p, o, dec0, (db, cb, scaling) = _case(7, 400)
    # the script-facing API (attas_sp_ml.py:153-159)
    x_gpu, info_gpu = fit.solve(p, dec0, db, cb, scaling, tol=1e-9)
    assert info_gpu['status'] == 'solved', info_gpu['status']
    s = nlp.InteriorPointSolver(OracleEvaluator(o), db, cb)
    s.add_num_option('tol', 1e-9)
    s.add_int_option('max_iter', 500)
    s.set_scaling(*scaling)
    x_cpu, info_cpu = s.solve(dec0)
    assert info_cpu['status'] == 'solved'
    assert info_gpu['iterations'] == info_cpu['iterations']
    vg, vc = p.variables(x_gpu), p.variables(x_cpu)
    for name in PARAMS:
        scale = max(1e-3, np.max(np.abs(vc[name])))
        np.testing.assert_allclose(vg[name], vc[name], rtol=1e-8,
                                   atol=1e-8 * scale, err_msg=name)
    np.testing.assert_allclose(info_gpu['obj'], info_cpu['obj'], rtol=1e-10)
    # the split the north star asks for: callbacks vs KKT factorisation
    assert info_gpu['seconds_kkt'] > 0 and info_gpu['seconds_callbacks'] > 0


def test_batch_fitter_matches_single_fits():
    cases = [_case(seed, 250) for seed in (3, 5, 11, 13)]
    problems = [c[1] for c in cases]
    db, cb, scaling = cases[0][4]
    singles = [fit.solve(c[1], c[3], *c[4], tol=1e-8, max_iter=400)
               for c in cases]
    bf = fit.BatchFitter(problems)
    batch = bf.fit([c[3] for c in cases], [c[4][0] for c in cases], cb,
                   scaling, tol=1e-8, max_iter=400)
    bf.close()
    for (xs, infos), (xb, infob) in zip(singles, batch):
        assert infos['status'] == infob['status'] == 'solved'
        assert infos['iterations'] == infob['iterations']
        np.testing.assert_allclose(xb, xs, rtol=1e-9, atol=1e-11)
    assert bf.launches < sum(i['callback_calls'] for _, i in singles)


def test_ipopt_callback_grouping_against_fake_ipopt(tmp_path):
    """The ctypes IPOPT binding driven by the CUDA evaluator: the fake
    libipopt (tests/fake_ipopt.c) calls every callback once; values must be
    the oracle's and the five callbacks must cost three kernel groups."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    lib = str(tmp_path / 'libipopt_fake.so')
    subprocess.run(['gcc', '-shared', '-fPIC', '-O1', '-o', lib,
                    os.path.join(here, 'fake_ipopt.c')], check=True)
    exp, p, o, dec0, (db, cb, scaling) = _case(3, 300)
    ev = nlp.GpuEvaluator(p)
    s = nlp.IpoptSolver(ev, db, cb, libpath=lib)
    s.set_scaling(*scaling)
    x, info = s.solve(dec0)
    assert info['status'] == 0 and info['last_error'] is None
    lam = 0.5 + 0.01 * np.arange(o.ncons)
    np.testing.assert_allclose(info['obj'], o.obj(dec0), rtol=1e-12)
    np.testing.assert_allclose(info['g'], o.constr(dec0), atol=1e-12)
    L = info['mult_x_L']
    np.testing.assert_allclose(L[0], o.constr_jac_val(dec0).sum(), rtol=1e-11)
    np.testing.assert_allclose(L[1], o.lag_hess_val(dec0, 0.75, lam).sum(),
                               rtol=1e-11)
    np.testing.assert_allclose(L[2], o.obj_grad(dec0).sum(), rtol=1e-11)
    assert ev.kernel_groups == 3
    s.close()


def test_parallel_batch_fitter_matches_thread_version():
    cases = [_case(seed, 250) for seed in (3, 5, 11, 13, 17)]
    problems = [c[1] for c in cases]
    db, cb, scaling = cases[0][4]
    dbs = [c[4][0] for c in cases]
    pf = fit.ParallelBatchFitter(problems, workers=3)
    par = pf.fit([c[3] for c in cases], dbs, cb, scaling, tol=1e-8,
                 max_iter=400)
    bf = fit.BatchFitter(problems)
    ser = bf.fit([c[3] for c in cases], dbs, cb, scaling, tol=1e-8,
                 max_iter=400)
    bf.close()
    for (xp, ip), (xs, is_) in zip(par, ser):
        assert ip['status'] == is_['status'] == 'solved'
        assert ip['iterations'] == is_['iterations']
        np.testing.assert_array_equal(xp, xs)
