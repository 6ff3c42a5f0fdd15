"""Generate the golden fixtures of tests/golden/ from the REFERENCE's own
unchanged ``symfem.py`` / ``fem.py`` hosted on the oracle engine.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Each fixture holds seeded inputs (y, u, dvec,
multipliers, obj_factor, dt), the recorded layout and every callback result
(objective, gradient, constraints, COO Jacobian, COO Lagrangian Hessian).
"""

import json
import os
import sys

import numpy as np

from oracle import refhost

#: (kind, nx, nu, ny, N) -- the three model sizes of SURVEY.md section 8 plus
#: every class composition the reference scripts build.
CASES = [
    ('innovation', 2, 1, 2, 7),
    ('balanced', 2, 1, 2, 6),
    ('ml', 2, 1, 2, 6),
    ('ml_zoh', 2, 1, 2, 5),
    ('ndisc_zoh', 2, 1, 2, 5),
    ('ml_balanced', 2, 1, 2, 5),
    ('innovation', 5, 3, 3, 9),
    ('balanced', 5, 3, 3, 5),
    ('ml_balanced', 5, 3, 3, 4),
    ('innovation', 4, 2, 7, 6),
    ('ndisc_zoh', 4, 2, 7, 3),
    ('innovation', 1, 1, 1, 2),
    ('innovation', 3, 2, 1, 70),
]


def seeded_point(problem, rng):
    """A generic evaluation point with positive sqrt-covariance diagonals."""
    dvec = rng.normal(size=problem.ndec)
    for name, spec in problem.decision.items():
        if name.endswith('_tril'):
            k = spec.size
            n = int(round((np.sqrt(8 * k + 1) - 1) / 2))
            diag = np.array([i == j for i in range(n) for j in range(i + 1)])
            spec.unpack_from(dvec)[diag] = rng.uniform(0.5, 2.0, size=n)
    return dvec


def layout_of(problem):
    return {
        'decision': [[n, list(s.shape), s.offset]
                     for n, s in problem.decision.items()],
        'dependent': [[n, list(s.shape), s.offset]
                      for n, s in problem.dependent.items()],
        'constraints': [[n, list(r.spec.shape), r.spec.offset]
                        for n, r in problem.constraints.items()],
        'objectives': [[n, list(r.spec.shape)]
                       for n, r in problem.objectives.items()],
    }


def generate(case, seed):
    kind, nx, nu, ny, N = case
    rng = np.random.default_rng(seed)
    y = rng.normal(size=(N, ny))
    u = rng.normal(size=(N, nu))
    dt = 0.05
    problem = refhost.make_problem(kind, y, u, nx, dt=dt)
    dvec = seeded_point(problem, rng)
    lam = rng.normal(size=problem.ncons)
    obj_factor = float(rng.uniform(0.5, 1.5))
    jr, jc = problem.constr_jac_ind()
    hr, hc = problem.lag_hess_ind()
    return dict(
        kind=kind, dims=np.array([nx, nu, ny]), N=N, dt=dt, y=y, u=u,
        dvec=dvec, lam=lam, obj_factor=obj_factor,
        layout=json.dumps(layout_of(problem)),
        ndec=problem.ndec, ncons=problem.ncons,
        f=problem.obj(dvec), grad=problem.obj_grad(dvec),
        g=problem.constr(dvec),
        jac_row=jr, jac_col=jc, jac_val=problem.constr_jac_val(dvec),
        hess_row=hr, hess_col=hc,
        hess_val=problem.lag_hess_val(dvec, obj_factor, lam))


def main(outdir=None):
    if not refhost.available():
        sys.exit('needs the reference tree (build container only)')
    here = os.path.dirname(os.path.abspath(__file__))
    outdir = outdir or os.path.join(os.path.dirname(here), 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    for seed, case in enumerate(CASES):
        kind, nx, nu, ny, N = case
        name = f'{kind}_nx{nx}_nu{nu}_ny{ny}_N{N}.npz'
        data = generate(case, 1000 + seed)
        np.savez_compressed(os.path.join(outdir, name), **data)
        print(name, 'ndec', data['ndec'], 'ncons', data['ncons'],
              'nnzjac', len(data['jac_val']), 'nnzhess', len(data['hess_val']))


if __name__ == '__main__':
    main()
