"""CPU checks of the parity-test machinery itself: the bit-exactness masks
(which Jacobian / Hessian entries are copies, negations or constants) and the
comparison helper, exercised with the oracle against the golden fixtures."""

import numpy as np

from colloc_fem_code_b200 import families
from oracle import ref_models
from parity_helpers import check_against, exact_masks, scale_of


def test_per_sample_derivatives_are_all_copy_type():
    """symfem.py:50-59 is bilinear: every per-sample Jacobian entry is +-1, a
    negated parameter or a negated state / input, every per-sample Hessian
    entry a constant times a multiplier.  Only d2 loglikelihood / d sRp_ii^2
    (1 / sRp_ii^2, N duplicates) needs a tolerance."""
    nx, nu, ny, N = 2, 1, 2, 50
    p = families.make_problem('innovation', np.zeros((N, ny)),
                              np.zeros((N, nu)), nx)
    jm, hm = exact_masks(p)
    assert jm.all()
    assert jm.size == (N - 1) * 20 + N * 18
    assert (~hm).sum() == N * ny            # the (sRp_ii, sRp_ii) duplicates
    assert hm.size == N * (2 * ny + 7) + (N - 1) * 8


def test_parameter_only_entries_split():
    nx, nu, ny, N = 2, 1, 2, 20
    p = families.make_problem('ml_zoh', np.zeros((N, ny)), np.zeros((N, nu)),
                              nx, dt=0.1)
    jm, hm = exact_masks(p)
    per_sample = (N - 1) * 20 + N * 18
    assert jm[:per_sample].all()
    # the ZOH discretisation is a cubic in Ac*dt: not copy-type
    assert not jm[per_sample:].all() and jm[per_sample:].any()
    assert not hm.all()


def test_oracle_matches_golden_with_bitwise_masks(golden):
    g = golden
    nx, nu, ny = g['dims']
    p = families.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])
    o = ref_models.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])
    d, s, lam = g['dvec'], g['obj_factor'], g['lam']
    res = {'f': o.obj(d), 'grad': o.obj_grad(d), 'g': o.constr(d),
           'jac': o.constr_jac_val(d), 'hess': o.lag_hess_val(d, s, lam)}
    ref = {'f': g['f'], 'grad': g['grad'], 'g': g['g'], 'jac': g['jac_val'],
           'hess': g['hess_val']}
    scale = scale_of(d, g['y'], g['u']) ** 2 * (nx + nu + ny + 1)
    check_against(res, ref, scale, exact_masks(p))
