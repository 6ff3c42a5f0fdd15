"""Synthetic trajectories and evaluation points (no datasets ship with the
reference: its ``data/`` directory is git-ignored).

The recipe restates the reference's Monte-Carlo data generator in NumPy
(/root/reference/mc_data_gen.m:5-16,46-54): a random stable discrete-time LTI
system driven by white inputs ``u ~ N(0, std_u^2)`` and process noise
``w ~ N(0, std_w^2)`` entering every state, measured with noise
``v ~ N(0, std_v^2)``.  MATLAB's ``drss`` is not reproducible outside MATLAB;
the stand-in draws eigenvalues inside a disc of radius ``rho`` (real ones and
complex pairs) and a random similarity transform.
"""

import numpy as np


def random_stable_system(rng, nx, nu, ny, rho=0.95):
    """Random (A, B, C, D) with spectral radius < rho (stand-in for drss)."""
    blocks = []
    i = 0
    while i < nx:
        if nx - i >= 2 and rng.random() < 0.5:
            r = rho * np.sqrt(rng.random())
            th = rng.uniform(0, np.pi)
            c, s = r * np.cos(th), r * np.sin(th)
            blocks.append(np.array([[c, s], [-s, c]]))
            i += 2
        else:
            blocks.append(np.array([[rng.uniform(-rho, rho)]]))
            i += 1
    Ad = np.zeros((nx, nx))
    k = 0
    for b in blocks:
        n = len(b)
        Ad[k:k + n, k:k + n] = b
        k += n
    T, _ = np.linalg.qr(rng.normal(size=(nx, nx)))
    T = T @ np.diag(rng.uniform(0.5, 2.0, size=nx))
    A = T @ Ad @ np.linalg.inv(T)
    B = rng.normal(size=(nx, nu))
    C = rng.normal(size=(ny, nx))
    D = rng.normal(size=(ny, nu))
    return A, B, C, D


def simulate(rng, N, A, B, C, D, std_u=1.0, std_w=0.05, std_v=0.2):
    """(u, y, x) of one experiment (mc_data_gen.m:49-54)."""
    nx, nu = B.shape
    ny = C.shape[0]
    u = std_u * rng.normal(size=(N, nu))
    w = std_w * rng.normal(size=(N, nx))
    v = std_v * rng.normal(size=(N, ny))
    x = np.empty((N, nx))
    x[0] = 0.0
    drive = u @ B.T + w
    if N > 4096:
        _simulate_blocked(A, drive, x)
    else:
        for k in range(N - 1):
            x[k + 1] = A @ x[k] + drive[k]
    y = x @ C.T + u @ D.T + v
    return u, y, x


def _simulate_blocked(A, drive, x):
    """Modal propagation for long records: with A = V diag(lam) V^-1 every
    mode is a first-order recursion, run by ``scipy.signal.lfilter``."""
    import scipy.signal
    lam, V = np.linalg.eig(A)
    zdrive = drive[:-1] @ np.linalg.inv(V).T
    z = np.zeros((len(x), len(lam)), dtype=complex)
    for i, li in enumerate(lam):
        z[1:, i] = scipy.signal.lfilter([1.0], [1.0, -li], zdrive[:, i])
    x[...] = (z @ V.T).real


def experiment(seed, N, nx, nu, ny, **kwargs):
    """Dict with the system and one simulated data record."""
    rng = np.random.default_rng(seed)
    A, B, C, D = random_stable_system(rng, nx, nu, ny)
    u, y, x = simulate(rng, N, A, B, C, D, **kwargs)
    return {'A': A, 'B': B, 'C': C, 'D': D, 'u': u, 'y': y, 'x': x}


def _ntril_dim(k):
    return int(round((np.sqrt(8 * k + 1) - 1) / 2))


def evaluation_point(problem, exp, seed=0):
    """Decision vector, multipliers and objective factor for benchmarking.

    Parameters are the true system perturbed by 1 %, the states are the
    simulated ones plus N(0, 1e-2), ``en ~ N(0, 1)``; every ``*_tril``
    square-root factor gets a diagonal in [0.5, 2]; all remaining
    (class-specific) decisions are N(0, 1).  ``lambda ~ N(0, 1)``.
    """
    rng = np.random.default_rng(seed + 12345)
    dvec = rng.normal(size=problem.ndec)
    var = problem.variables(dvec)
    for name in ('A', 'B', 'C', 'D'):
        if name in problem.decision:
            var[name][...] = exp[name] * (1 + 0.01 * rng.normal(
                size=exp[name].shape))
    var['ybias'][...] = 0.01 * rng.normal(size=var['ybias'].shape)
    if 'Ln' in problem.decision:
        var['Ln'][...] = 0.1 * rng.normal(size=var['Ln'].shape)
    var['x'][...] = exp['x'] + 0.1 * rng.normal(size=exp['x'].shape)
    for name, spec in problem.decision.items():
        if name.endswith('_tril'):
            n = _ntril_dim(spec.size)
            tri = var[name]
            tri[...] = 0.1 * rng.normal(size=spec.size)
            rows, cols = np.tril_indices(n)
            tri[rows == cols] = rng.uniform(0.5, 2.0, size=n)
    lam = rng.normal(size=problem.ncons)
    return dvec, lam, 1.0
