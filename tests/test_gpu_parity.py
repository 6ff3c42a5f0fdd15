"""Parity of the CUDA path (through the C ABI) with the CPU oracle.

Tolerances (FP64): 1e-12 relative, as BASELINE.json's north star states.  For
constraint values, which are differences of O(1) terms that cancel at a
solution, "relative" is relative to the magnitude of the terms
(|value| + scale), not to the possibly tiny residual.
"""

import numpy as np
import pytest

from colloc_fem_code_b200 import backend, families, synthetic
from oracle import ref_models

pytestmark = pytest.mark.gpu

from parity_helpers import (RTOL, check_against, exact_masks,
                            grad_sample_mask)
from parity_helpers import scale_of as _scale

MASKS = (1, 3, 4, 8, 16, 5, 15, 31)


def cuda_callbacks(p, dvec, sigma, lam):
    """The five callbacks, one C-ABI call each (the IPOPT call pattern)."""
    return {'f': p.obj(dvec), 'grad': p.obj_grad(dvec), 'g': p.constr(dvec),
            'jac': p.constr_jac_val(dvec),
            'hess': p.lag_hess_val(dvec, sigma, lam)}


def oracle_callbacks(o, dvec, sigma, lam):
    return {'f': o.obj(dvec), 'grad': o.obj_grad(dvec), 'g': o.constr(dvec),
            'jac': o.constr_jac_val(dvec),
            'hess': o.lag_hess_val(dvec, sigma, lam)}


def test_golden_fixtures(golden):
    g = golden
    nx, nu, ny = g['dims']
    p = families.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])
    ref = {'f': g['f'], 'grad': g['grad'], 'g': g['g'], 'jac': g['jac_val'],
           'hess': g['hess_val']}
    scale = _scale(g['dvec'], g['y'], g['u']) ** 2 * (nx + nu + ny + 1)
    exact = exact_masks(p)
    check_against(cuda_callbacks(p, g['dvec'], g['obj_factor'], g['lam']),
                  ref, scale, exact)
    # fused single pass
    f, grad, gg, jac, hess = p.backend.eval_all(g['dvec'], g['obj_factor'],
                                                g['lam'])
    check_against({'f': f, 'grad': grad, 'g': gg, 'jac': jac, 'hess': hess},
                  ref, scale, exact)


def test_every_kernel_variant(golden):
    """Each instantiated mask produces the same bits as the full kernel."""
    g = golden
    if g['kind'] not in ('innovation', 'ml_balanced', 'ndisc_zoh'):
        pytest.skip('variant sweep runs on three families')
    nx, nu, ny = g['dims']
    p = families.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])
    h = p.backend.handle
    h.set_dvec(g['dvec'])
    h.set_multipliers(g['obj_factor'], g['lam'])
    h.eval(backend.ALL)
    full = {b: h.fetch(b).copy() for b in (1, 2, 4, 8, 16)}
    for mask in MASKS:
        h.set_dvec(g['dvec'])      # invalidates cached results
        h.eval(mask)
        for b in (1, 2, 4, 8, 16):
            if mask & b:
                np.testing.assert_array_equal(h.fetch(b), full[b],
                                              err_msg=f'mask {mask} bit {b}')


CASES = [
    ('innovation', (2, 1, 2), 2), ('innovation', (2, 1, 2), 3),
    ('innovation', (2, 1, 2), 257), ('innovation', (2, 1, 2), 10000),
    ('balanced', (2, 1, 2), 1000), ('ml', (2, 1, 2), 1000),
    ('ml_zoh', (2, 1, 2), 300), ('ndisc_zoh', (2, 1, 2), 257),
    ('innovation', (5, 3, 3), 257), ('balanced', (5, 3, 3), 250),
    ('ml_balanced', (5, 3, 3), 250), ('innovation', (5, 3, 3), 10000),
    ('innovation', (4, 2, 7), 2), ('innovation', (4, 2, 7), 129),
    ('ndisc_zoh', (4, 2, 7), 601), ('innovation', (4, 2, 7), 10000),
    ('innovation', (1, 1, 1), 33), ('innovation', (3, 2, 1), 4097),
    # extension without reference counterpart: trapezoidal collocation
    ('trapezoid', (2, 1, 2), 2), ('trapezoid', (2, 1, 2), 1000),
    ('trapezoid', (3, 2, 2), 4097),
    # BASELINE sizes: the exact bench.py workload (attas_sp_ml at N = 1e6) and
    # the other two script shapes, entry for entry against the oracle
    ('ml', (2, 1, 2), 1_000_000), ('balanced', (5, 3, 3), 1_000_000),
    ('ndisc_zoh', (4, 2, 7), 1_000_000),
]


@pytest.mark.parametrize('kind,dims,N', CASES,
                         ids=[f'{k}-{d[0]}{d[1]}{d[2]}-N{n}'
                              for k, d, n in CASES])
def test_random_point_vs_oracle(kind, dims, N):
    nx, nu, ny = dims
    exp = synthetic.experiment(N + 17, N, nx, nu, ny)
    p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
    o = ref_models.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
    dvec, lam, _ = synthetic.evaluation_point(p, exp, seed=N)
    sigma = 0.8
    scale = _scale(dvec, exp['y'], exp['u']) ** 2 * (nx + nu + ny + 1)
    res = cuda_callbacks(p, dvec, sigma, lam)
    check_against(res, oracle_callbacks(o, dvec, sigma, lam), scale,
                  exact_masks(p), (grad_sample_mask(p), N))
    # the summed parameter entries against their closed form (adfem.py:119:
    # sum over samples of -1 / sRp_ii), at the stated 1e-12
    if kind != 'trapezoid':
        var = p.variables(dvec)
        diag = var['sRp_tril'][families.models.tril_diag(ny)]
        sl = p.decision['sRp_tril']
        got = res['grad'][sl.offset:sl.offset + sl.size][
            families.models.tril_diag(ny)]
        np.testing.assert_allclose(got, -N / diag, rtol=RTOL)
        f_ref = -0.5 * np.sum(var['en'] ** 2) - N * np.log(diag).sum()
        np.testing.assert_allclose(res['f'], f_ref, rtol=RTOL)


def test_bench_workload_vs_oracle():
    """bench.py's own inputs (same family, dims, seed, N and evaluation point
    as the timed run), fused single pass, against the oracle."""
    import bench
    nx, nu, ny = bench.DIMS
    N = bench.N_PER_GPU
    exp = synthetic.experiment(0, N, nx, nu, ny)
    p = families.make_problem(bench.KIND, exp['y'], exp['u'], nx)
    o = ref_models.make_problem(bench.KIND, exp['y'], exp['u'], nx)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    f, grad, g, jac, hess = p.backend.eval_all(dvec, sigma, lam)
    scale = _scale(dvec, exp['y'], exp['u']) ** 2 * (nx + nu + ny + 1)
    check_against({'f': f, 'grad': grad, 'g': g, 'jac': jac, 'hess': hess},
                  oracle_callbacks(o, dvec, sigma, lam), scale,
                  exact_masks(p), (grad_sample_mask(p), N))


def test_known_answer_noise_free():
    """True states of noise-free data with en = 0: all defects are 0 to
    rounding and f = -N log det sRp (symfem.py:50-65)."""
    nx, nu, ny, N = 4, 2, 3, 5000
    exp = synthetic.experiment(3, N, nx, nu, ny, std_w=0.0, std_v=0.0)
    p = families.make_problem('innovation', exp['y'], exp['u'], nx)
    dvec = np.zeros(p.ndec)
    var = p.variables(dvec)
    for k in 'ABCD':
        var[k][...] = exp[k]
    var['x'][...] = exp['x']
    diag = np.array([1.5, 0.7, 1.1])
    var['sRp_tril'][families.models.tril_diag(ny)] = diag
    g = p.constr(dvec)
    assert np.max(np.abs(g)) <= 1e-12 * _scale(exp['x'], exp['y'])
    np.testing.assert_allclose(p.obj(dvec), -N * np.log(diag).sum(),
                               rtol=RTOL)
    grad = p.obj_grad(dvec)
    sl = p.decision['sRp_tril']
    np.testing.assert_allclose(
        grad[sl.offset:sl.offset + sl.size][families.models.tril_diag(ny)],
        -N / diag, rtol=RTOL)


@pytest.mark.parametrize('dims', [(2, 1, 2), (5, 3, 3)],
                         ids=['attas212', 'blackbox533'])
def test_full_size_million_samples(dims):
    """BASELINE size N = 1e6: direct NumPy restatement of the per-sample
    values, and size-independent properties of the derivative values: the
    per-sample functions are bilinear, so central differences are exact,
    J(x) v = (g(x+v) - g(x-v)) / 2 and H(lam) v = (J(x+v)' - J(x-v)') lam / 2,
    and the Hessian is linear in (sigma, lambda)."""
    import scipy.sparse as sp
    nx, nu, ny = dims
    N = 1_000_000
    exp = synthetic.experiment(0, N, nx, nu, ny)
    p = families.make_problem('innovation', exp['y'], exp['u'], nx)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    var = p.variables(dvec)
    x, en, y, u = var['x'], var['en'], exp['y'], exp['u']
    sRp = families.models.tril_mat(var['sRp_tril'])
    dyn = x[1:] - (x[:-1] @ var['A'].T + u[:-1] @ var['B'].T
                   + en[:-1] @ var['Ln'].T)
    inn = y - (x @ var['C'].T + u @ var['D'].T + var['ybias']) - en @ sRp.T
    f = -0.5 * np.sum(en ** 2) - N * np.log(np.diag(sRp)).sum()
    g = p.constr(dvec)
    scale = _scale(dvec, y, u) ** 2 * (nx + nu + ny + 1)
    np.testing.assert_allclose(g[:dyn.size], dyn.ravel(), rtol=RTOL,
                               atol=RTOL * scale)
    np.testing.assert_allclose(g[dyn.size:], inn.ravel(), rtol=RTOL,
                               atol=RTOL * scale)
    np.testing.assert_allclose(p.obj(dvec), f, rtol=RTOL)
    grad = p.obj_grad(dvec)
    eo = p.decision['en'].offset
    np.testing.assert_array_equal(grad[eo:eo + en.size], -en.ravel())

    rng = np.random.default_rng(9)
    v = rng.normal(size=p.ndec)
    jr, jc = p.constr_jac_ind()
    jac = p.constr_jac_val(dvec)
    J = sp.csr_matrix((jac, (jr, jc)), shape=(p.ncons, p.ndec))
    fd = 0.5 * (p.constr(dvec + v) - p.constr(dvec - v))
    np.testing.assert_allclose(J @ v, fd, rtol=1e-11,
                               atol=1e-11 * scale * _scale(v))
    hr, hc = p.lag_hess_ind()
    hess0 = p.lag_hess_val(dvec, 0.0, lam)       # constraint part only
    H = sp.csr_matrix((hess0, (hr, hc)), shape=(p.ndec, p.ndec))
    H = H + sp.tril(H, -1).T
    Jp = sp.csr_matrix((p.constr_jac_val(dvec + v), (jr, jc)),
                       shape=(p.ncons, p.ndec))
    Jm = sp.csr_matrix((p.constr_jac_val(dvec - v), (jr, jc)),
                       shape=(p.ncons, p.ndec))
    fdh = 0.5 * ((Jp.T @ lam) - (Jm.T @ lam))
    np.testing.assert_allclose(H @ v, fdh, rtol=1e-10,
                               atol=1e-10 * scale * _scale(v))
    # linearity in the multipliers
    hess1 = p.lag_hess_val(dvec, sigma, lam)
    hess2 = p.lag_hess_val(dvec, 2 * sigma, 2 * lam)
    np.testing.assert_allclose(hess2, 2 * hess1, rtol=1e-15)


@pytest.mark.parametrize('kind,dims,N,world', [
    ('innovation', (2, 1, 2), 1001, 2), ('ml_balanced', (2, 1, 2), 777, 3),
    ('ndisc_zoh', (4, 2, 7), 301, 2), ('innovation', (5, 3, 3), 4099, 4),
    ('trapezoid', (3, 2, 2), 1025, 3)])
def test_time_shards_on_one_gpu(kind, dims, N, world):
    """The halo = 1 handles of a time-sharded trajectory, evaluated one after
    the other on ONE GPU: the assembled results equal the unsharded CUDA
    evaluation bit for bit (g, Jacobian, Hessian, per-sample gradient) and the
    summed [f, d f/d params] partials equal it to rounding."""
    from colloc_fem_code_b200 import sharding
    nx, nu, ny = dims
    exp = synthetic.experiment(N, N, nx, nu, ny)
    p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp, seed=1)
    ref = cuda_callbacks(p, dvec, sigma, lam)
    full = {k: np.zeros_like(ref[k]) for k in ('grad', 'g', 'jac', 'hess')}
    red_sum = None
    for rank in range(world):
        ev = sharding.ShardedEvaluator(p, rank, world)
        ev.set_point(dvec, sigma, lam)
        h = ev.handle
        h.eval(backend.ALL)
        loc = {k: h.fetch(b) for k, b in (('grad', 2), ('g', 4), ('jac', 8),
                                          ('hess', 16))}
        for k in full:
            ev.shard.scatter(k, loc[k], full[k])
        import ctypes
        red = np.empty(ev.n_reduce)
        # the [n_reduce] partial of this shard (device buffer -> host)
        import torch
        ptr = h.device_ptrs()['reduce']

        class Arr:
            __cuda_array_interface__ = {'shape': (ev.n_reduce,),
                                        'typestr': '<f8',
                                        'data': (ptr, False), 'version': 2}
        red[:] = torch.as_tensor(Arr(), device='cuda:0').cpu().numpy()
        red_sum = red if red_sum is None else red_sum + red
        h.close()
    for k in ('g', 'jac', 'hess'):
        np.testing.assert_array_equal(full[k], ref[k], err_msg=k)
    np.testing.assert_allclose(red_sum[0], ref['f'], rtol=1e-13)
    # parameter entries of the gradient come from the summed partials
    slots = ev.lib.model['reduce'][1:]
    st = p.structure
    for (var, flat), val in zip(slots, red_sum[1:]):
        idx = p.decision[st.var_names[var]].offset + flat
        full['grad'][idx] = val
    np.testing.assert_allclose(full['grad'], ref['grad'], rtol=1e-13,
                               atol=1e-300)


def test_batched_problems_match_single_problem_evaluation():
    """cfem_create(batch = B): blockIdx.y = problem (Monte-Carlo batches)."""
    nx, nu, ny, N, B = 5, 3, 3, 250, 5
    cases = []
    for b in range(B):
        exp = synthetic.experiment(100 + b, N, nx, nu, ny)
        p = families.make_problem('ml_balanced', exp['y'], exp['u'], nx)
        cases.append((p,) + synthetic.evaluation_point(p, exp, seed=b))
    p0 = cases[0][0]
    st = p0.structure
    lib = backend.Library.for_structure(st)
    data = [np.stack([c[0].structure.data[i]['source'] for c in cases])
            for i in range(len(st.data))]
    h = backend.Handle(lib, N, data, st.scalar_values, batch=B)
    h.set_dvec(np.stack([c[1] for c in cases]))
    h.set_multipliers(0.9, np.stack([c[2] for c in cases]))
    h.eval(backend.ALL)
    out = {k: h.fetch(b) for k, b in (('f', 1), ('grad', 2), ('g', 4),
                                      ('jac', 8), ('hess', 16))}
    for b, (p, dvec, lam, _) in enumerate(cases):
        one = p.backend.eval_all(dvec, 0.9, lam)
        for k, v in zip(('f', 'grad', 'g', 'jac', 'hess'), one):
            np.testing.assert_array_equal(np.ravel(out[k][b]), np.ravel(v),
                                          err_msg=f'{k} problem {b}')


def test_reductions_are_bitwise_reproducible():
    """No floating-point atomics, fixed summation tree: repeated launches
    (different CTA scheduling, cold and warm caches) give identical bits for
    the objective and the whole gradient."""
    nx, nu, ny, N = 2, 1, 2, 300_000
    exp = synthetic.experiment(4, N, nx, nu, ny)
    p = families.make_problem('ml', exp['y'], exp['u'], nx)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    h = p.backend.handle
    seen = set()
    for rep in range(6):
        if rep % 2:
            h.flush_l2(256 << 20)
        h.set_dvec(dvec)
        h.set_multipliers(sigma, lam)
        h.eval(backend.ALL if rep < 3 else backend.F | backend.GRAD)
        f = h.fetch(backend.F).tobytes()
        grad = h.fetch(backend.GRAD).tobytes()
        seen.add((f, grad))
    assert len(seen) == 1


@pytest.mark.parametrize('kind,dims,N,batch', [
    ('ml_balanced', (5, 3, 3), 250, 3), ('ndisc_zoh', (4, 2, 7), 601, 1),
    ('ml', (2, 1, 2), 70_001, 1)])
def test_launch_variants_give_identical_bits(kind, dims, N, batch, monkeypatch):
    """The two kernels of a callback set overlapped by programmatic dependent
    launch on one stream (default), forked/joined over an auxiliary stream
    (CFEM_PDL=0) and the CUDA-graph form of the latter (CFEM_GRAPH=1) are the
    same arithmetic: results must agree bit for bit, for every callback
    subset."""
    nx, nu, ny = dims
    cases = []
    for b in range(batch):
        exp = synthetic.experiment(40 + b, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
        cases.append((p,) + synthetic.evaluation_point(p, exp, seed=b))
    st = cases[0][0].structure
    lib = backend.Library.for_structure(st)
    data = [np.stack([c[0].structure.data[i]['source'] for c in cases])
            for i in range(len(st.data))]
    if batch == 1:
        data = [d[0] for d in data]
    dvec = np.stack([c[1] for c in cases])
    lam = np.stack([c[2] for c in cases])
    results = {}
    for label, env in (('pdl', {'CFEM_PDL': '2'}),
                       ('fork-join', {'CFEM_PDL': '0'}),
                       ('graph', {'CFEM_PDL': '0', 'CFEM_GRAPH': '1'})):
        for k in ('CFEM_PDL', 'CFEM_GRAPH'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h = backend.Handle(lib, N, data, st.scalar_values, batch=batch)
        got = []
        for rep, mask in enumerate((backend.ALL, backend.F | backend.G,
                                    backend.HESS, backend.JAC, backend.ALL)):
            h.set_dvec(dvec * (1 + 1e-3 * rep))
            h.set_multipliers(0.5 + rep, lam)       # new kernel arguments
            h.eval(mask)
            for bit in (1, 2, 4, 8, 16):
                if mask & bit:
                    got.append(h.fetch(bit).copy())
        results[label] = got
        h.close()
    for label in ('fork-join', 'graph'):
        for a, b in zip(results['pdl'], results[label]):
            np.testing.assert_array_equal(a, b, err_msg=label)
    # and against the oracle-checked single-problem path
    p, d0, l0, _ = cases[0]
    one = p.backend.eval_all(d0 * (1 + 4e-3), 4.5, l0)
    last = results['pdl'][-5:]
    for a, b in zip(last, one):
        np.testing.assert_array_equal(np.ravel(np.asarray(a))[:np.size(b)]
                                      if batch > 1 else np.ravel(a),
                                      np.ravel(b))


@pytest.mark.parametrize('batch', [1, 3])
def test_single_copy_transfers(batch):
    """cfem_set_inputs / cfem_fetch_results_async (one copy per direction
    through host blocks that mirror the device slabs) against the per-array
    calls, for full and partial result masks."""
    nx, nu, ny, N = 2, 1, 2, 1001
    cases = []
    for b in range(batch):
        exp = synthetic.experiment(60 + b, N, nx, nu, ny)
        p = families.make_problem('ml', exp['y'], exp['u'], nx)
        cases.append((p,) + synthetic.evaluation_point(p, exp, seed=b))
    st = cases[0][0].structure
    lib = backend.Library.for_structure(st)
    data = [np.stack([c[0].structure.data[i]['source'] for c in cases])
            for i in range(len(st.data))]
    if batch == 1:
        data = [d[0] for d in data]
    h = backend.Handle(lib, N, data, st.scalar_values, batch=batch)
    dvec = np.concatenate([c[1] for c in cases])
    lam = np.concatenate([c[2] for c in cases])
    in_off, in_total, res_off, res_total = h.io_layout()
    assert all(o % 32 == 0 for o in in_off + res_off)
    buf = backend.HostBuffers(h)
    assert buf.inputs.size == in_total and buf.results.size == res_total
    names = ('f', 'grad', 'g', 'jac', 'hess')
    # reference: per-array calls
    h.set_dvec(dvec)
    h.set_multipliers(0.75, lam)
    h.eval(backend.ALL)
    ref = {n: h.fetch(1 << i).copy() for i, n in enumerate(names)}
    # one copy each way
    buf.dvec[:] = dvec
    buf.lam[:] = lam
    buf.results[:] = np.nan
    buf.upload(0.75)
    h.eval(backend.ALL)
    buf.fetch_all()
    for n in names:
        np.testing.assert_array_equal(getattr(buf, n), np.ravel(ref[n]), n)
    # the overlapped whole-set call: lambda goes up while the first results
    # come down (cfem_eval_callback_set)
    buf.dvec[:] = dvec
    buf.lam[:] = lam
    buf.results[:] = np.nan
    buf.callback_set(0.75)
    for n in names:
        np.testing.assert_array_equal(getattr(buf, n), np.ravel(ref[n]), n)
    buf.lam[:] = 2.0 * lam
    buf.callback_set(1.5)
    np.testing.assert_array_equal(buf.hess, 2.0 * np.ravel(ref['hess']))
    np.testing.assert_array_equal(buf.jac, np.ravel(ref['jac']))
    buf.lam[:] = lam
    # x only, partial masks: the copied range spans first..last requested
    buf.dvec[:] = dvec * 1.001
    buf.upload()
    h.eval(backend.F | backend.G)
    buf.results[:] = np.nan
    buf.fetch(backend.F | backend.G)
    h.set_dvec(dvec * 1.001)
    h.eval(backend.F | backend.G)
    np.testing.assert_array_equal(buf.f, np.ravel(h.fetch(backend.F)))
    np.testing.assert_array_equal(buf.g, np.ravel(h.fetch(backend.G)))
    assert np.isnan(buf.jac).all() and np.isnan(buf.hess).all()
    with pytest.raises(backend.CfemError):
        buf.fetch(backend.HESS)             # not evaluated at this x
    buf.close()
    h.close()
