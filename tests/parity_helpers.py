"""Shared pieces of the parity tests (CUDA path vs CPU oracle).

Tolerances (FP64): 1e-12 relative, as BASELINE.json's north star states, for
everything that involves more than one rounding (objective, reduced gradient
entries, constraint values, polynomial parameter-only entries).  Jacobian and
Hessian entries whose generated expression is a copy, a negation or a constant
(-A_ij, -x_kj, +-1, +-lambda: /root/reference/symfem.py:50-59 is bilinear, so
that is every per-sample derivative entry) involve at most ONE rounding and
must be bit-identical to the oracle's.
"""

import re

import numpy as np

RTOL = 1e-12

_NUM = r'[0-9]+(?:\.[0-9]*)?(?:[eE][-+]?[0-9]+)?'
_IDENT = r'v_[A-Za-z0-9_]+'
#: [-] number | [-] ident | [-] number*ident : at most one rounding
_JAC_EXACT = re.compile(rf'^\s*-?\s*(?:{_NUM}|{_IDENT}|{_NUM}\s*\*\s*{_IDENT})\s*$')
#: Hessian entries are multiplied by lambda_i / obj_factor: a constant second
#: derivative times the multiplier is one rounding
_HESS_EXACT = re.compile(rf'^\s*-?\s*{_NUM}\s*$')


def _mask(p, blocks, pattern, total):
    st = p.structure
    rows = st.fun_rows(st.N)
    out = np.zeros(total, dtype=bool)
    off = 0
    for blk in blocks:
        M, c = rows[blk['fun']], blk['c']
        ex = np.array([bool(pattern.match(e.code)) for e in blk['entries']])
        out[off:off + M * c].reshape(M, c)[:] = ex
        off += M * c
    assert off == total, (off, total)
    return out


def exact_masks(p):
    """Boolean masks over the Jacobian / Hessian COO value arrays of problem
    ``p``: True where the generated expression involves at most one rounding
    (copy, negation, constant, constant times multiplier)."""
    st = p.structure
    return (_mask(p, st.jac_blocks, _JAC_EXACT, p.nnzjac),
            _mask(p, st.hess_blocks, _HESS_EXACT, p.nnzhess))


def grad_sample_mask(p):
    """True for the gradient entries of per-sample variables (x, en): they are
    -en copies or structural zeros, i.e. bit-exact; the others are parameter
    entries summed over samples (adfem.py:119)."""
    st = p.structure
    out = np.zeros(p.ndec, dtype=bool)
    for v in st.vars:
        if v['per_sample']:
            d = p.decision[v['name']]
            out[d.offset:d.offset + d.size] = True
    return out


def scale_of(*arrays):
    return max(1.0, *(float(np.max(np.abs(a))) for a in arrays if a.size))


def check_against(res, ref, scale, exact=None, grad_sample=None):
    """res / ref: dicts with f, grad, g, jac, hess.  ``exact``: the pair of
    masks of :func:`exact_masks` -- those entries are compared bit for bit,
    the rest to 1e-12 relative.  ``grad_sample``: (:func:`grad_sample_mask`,
    number of samples) -- splits the gradient the same way."""
    np.testing.assert_allclose(res['f'], ref['f'], rtol=RTOL)
    if grad_sample is None:
        np.testing.assert_allclose(res['grad'], ref['grad'], rtol=RTOL,
                                   atol=1e-300)
    else:
        gs, n_rows = grad_sample
        np.testing.assert_array_equal(res['grad'][gs], ref['grad'][gs],
                                      err_msg='grad: per-sample block')
        # parameter entries are sums over samples.  The reference (and the
        # oracle) add them up sequentially -- reshape(-1, ...).sum(0),
        # adfem.py:119 -- with a rounding-error bound of n_rows * eps / 2
        # relative for same-sign terms; the CUDA tree sum is the accurate
        # one (checked against the closed form -N / sRp_ii at 1e-12 in
        # test_gpu_parity and in bench.py's reduce_check).
        rtol = RTOL + 0.5 * n_rows * np.finfo(float).eps
        np.testing.assert_allclose(res['grad'][~gs], ref['grad'][~gs],
                                   rtol=rtol, atol=1e-300)
    np.testing.assert_allclose(res['g'], ref['g'], rtol=RTOL,
                               atol=RTOL * scale)
    for key, mask in zip(('jac', 'hess'), exact or (None, None)):
        a, b = np.ravel(res[key]), np.ravel(ref[key])
        if mask is None:
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=RTOL * scale)
            continue
        np.testing.assert_array_equal(a[mask], b[mask],
                                      err_msg=f'{key}: copy-type entries')
        np.testing.assert_allclose(a[~mask], b[~mask], rtol=RTOL,
                                   atol=RTOL * scale)
