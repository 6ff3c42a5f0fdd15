"""Build-container only: the oracle restatement (oracle/ref_models.py) against
the reference's own UNCHANGED symfem.py / fem.py hosted on the oracle engine,
entry for entry, plus the golden fixtures being reproducible from it."""

import numpy as np
import pytest

from oracle import make_golden, ref_models, refhost

pytestmark = pytest.mark.skipif(
    not refhost.available(),
    reason='/root/reference is only present in the build container')

CASES = [('innovation', 2, 1, 2, 5), ('ml_balanced', 2, 1, 2, 4),
         ('ndisc_zoh', 3, 1, 2, 4), ('balanced', 3, 2, 2, 6),
         ('ml_zoh', 3, 2, 1, 4)]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(map(str, c)))
def test_restatement_equals_reference(case):
    kind, nx, nu, ny, N = case
    rng = np.random.default_rng(77)
    y = rng.normal(size=(N, ny))
    u = rng.normal(size=(N, nu))
    ref = refhost.make_problem(kind, y, u, nx, dt=0.02)
    ours = ref_models.make_problem(kind, y, u, nx, dt=0.02)
    assert make_golden.layout_of(ref) == make_golden.layout_of(ours)
    dvec = make_golden.seeded_point(ref, rng)
    lam = rng.normal(size=ref.ncons)
    for a, b in zip(ref.constr_jac_ind() + ref.lag_hess_ind(),
                    ours.constr_jac_ind() + ours.lag_hess_ind()):
        np.testing.assert_array_equal(a, b)
    kw = dict(rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ours.obj(dvec), ref.obj(dvec), rtol=1e-14)
    np.testing.assert_allclose(ours.obj_grad(dvec), ref.obj_grad(dvec), **kw)
    np.testing.assert_allclose(ours.constr(dvec), ref.constr(dvec),
                               rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(ours.constr_jac_val(dvec),
                               ref.constr_jac_val(dvec), **kw)
    np.testing.assert_allclose(ours.lag_hess_val(dvec, 0.9, lam),
                               ref.lag_hess_val(dvec, 0.9, lam), **kw)


def test_golden_fixture_is_reproducible():
    case = make_golden.CASES[0]
    data = make_golden.generate(case, 1000)
    import os
    from conftest import GOLDEN_DIR, load_golden
    kind, nx, nu, ny, N = case
    g = load_golden(os.path.join(
        GOLDEN_DIR, f'{kind}_nx{nx}_nu{nu}_ny{ny}_N{N}.npz'))
    np.testing.assert_array_equal(data['dvec'], g['dvec'])
    np.testing.assert_allclose(data['jac_val'], g['jac_val'], rtol=1e-15)
    np.testing.assert_allclose(data['hess_val'], g['hess_val'], rtol=1e-15)
