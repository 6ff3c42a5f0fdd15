"""High-precision (mpmath, 40 digits) evaluation of the oracle's symbolic
expressions on a golden fixture's inputs -> tests/golden/exact/*.npz.

TEST INFRASTRUCTURE ONLY.  Gives the FP64 results of both the NumPy oracle and
the CUDA path a common, rounding-free yardstick: the fixture stores the
correctly rounded values and, for every constraint value, the sum of the
magnitudes of its terms (the scale its cancellation error is relative to).

    python -m oracle.make_exact
"""

import os

import mpmath
import numpy as np
import sympy

from oracle import ref_models

CASES = ['innovation_nx2_nu1_ny2_N7', 'ml_balanced_nx2_nu1_ny2_N5',
         'ndisc_zoh_nx2_nu1_ny2_N5', 'innovation_nx4_nu2_ny7_N6']


def _mp_eval(fun, exprs, values, rows):
    """[rows, len(exprs)] high-precision values of ``exprs``."""
    flat_syms = [s for a in fun.args for s in fun.arg_syms[a].ravel()]
    f = sympy.lambdify(flat_syms, list(exprs), modules='mpmath')
    out = []
    for k in range(rows):
        args = []
        for a in fun.args:
            core = fun.arg_syms[a].shape
            val = np.asarray(values[a], dtype=float)
            ext = val.shape[:val.ndim - len(core)]
            flat = val.reshape(ext + (-1,))
            row = flat[k] if ext else flat
            args += [mpmath.mpf(float(v)) for v in np.atleast_1d(row)]
        out.append(f(*args))
    return out


def exact(golden_path):
    mpmath.mp.dps = 40
    g = np.load(golden_path)
    nx, nu, ny = (int(v) for v in g['dims'])
    p = ref_models.make_problem(str(g['kind']), g['y'], g['u'], nx,
                                dt=float(g['dt']))
    dvec = g['dvec']
    lam = g['lam']
    sigma = mpmath.mpf(float(g['obj_factor']))
    var = p.variables(dvec)
    f_total = mpmath.mpf(0)
    for reg in p.objectives.values():
        for row in _mp_eval(reg.fun, reg.fun.out.ravel(), var, reg.M):
            f_total += sum(row)
    gvals, gscale = [], []
    for reg in p.constraints.values():
        exprs = list(reg.fun.out.ravel())
        mags = [sum(sympy.Abs(t) for t in sympy.Add.make_args(
            sympy.expand(e))) for e in exprs]
        vals = _mp_eval(reg.fun, exprs, var, reg.M)
        scal = _mp_eval(reg.fun, mags, var, reg.M)
        gvals += [float(v) for row in vals for v in row]
        gscale += [float(v) for row in scal for v in row]
    jac = []
    for reg, wrt, entries in p._jac_blocks():
        rows = _mp_eval(reg.fun, [e[-1] for e in entries], var, reg.M)
        jac += [float(v) for row in rows for v in row]
    hess = []
    for reg, is_obj, pair, entries in p._hess_blocks():
        rows = _mp_eval(reg.fun, [e[-1] for e in entries], var, reg.M)
        if not is_obj:
            lam_f = reg.spec.unpack_from(lam).reshape(reg.M, reg.out_core)
        for k, row in enumerate(rows):
            for e, v in zip(entries, row):
                mult = sigma if is_obj else mpmath.mpf(float(lam_f[k, e[2]]))
                hess.append(float(mult * v))
    return dict(f=float(f_total), g=np.array(gvals), g_scale=np.array(gscale),
                jac_val=np.array(jac), hess_val=np.array(hess))


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    gold = os.path.join(os.path.dirname(here), 'tests', 'golden')
    os.makedirs(os.path.join(gold, 'exact'), exist_ok=True)
    for name in CASES:
        data = exact(os.path.join(gold, name + '.npz'))
        np.savez_compressed(os.path.join(gold, 'exact', name + '.npz'), **data)
        print(name, len(data['g']), len(data['jac_val']),
              len(data['hess_val']))


if __name__ == '__main__':
    main()
