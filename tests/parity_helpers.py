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


def scale_of(*arrays):
    return max(1.0, *(float(np.max(np.abs(a))) for a in arrays if a.size))


def check_against(res, ref, scale, exact=None):
    """res / ref: dicts with f, grad, g, jac, hess.  ``exact``: the pair of
    masks of :func:`exact_masks` -- those entries are compared bit for bit,
    the rest to 1e-12 relative."""
    np.testing.assert_allclose(res['f'], ref['f'], rtol=RTOL)
    np.testing.assert_allclose(res['grad'], ref['grad'], rtol=RTOL,
                               atol=1e-300)
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
