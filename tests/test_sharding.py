"""Time-sharding on the CPU: bookkeeping of
colloc_fem_code_b200.sharding, and a world_size-2 gloo run in which every rank
evaluates its shard with the CPU oracle (the product has no CPU evaluator),
all-reduces [objective, parameter gradient] and the assembled results are
compared with the unsharded oracle."""

import os
import socket

import numpy as np
import pytest

from colloc_fem_code_b200 import families, sharding, synthetic
from oracle import ref_models


def _setup(kind, dims, N, seed=3):
    nx, nu, ny = dims
    exp = synthetic.experiment(seed, N, nx, nu, ny)
    p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp, seed)
    return exp, p, dvec, lam, sigma


@pytest.mark.parametrize('world', [1, 2, 3, 8])
@pytest.mark.parametrize('kind', ['innovation', 'ml_balanced'])
def test_shards_tile_the_global_vectors(kind, world):
    exp, p, dvec, lam, sigma = _setup(kind, (2, 1, 2), 37)
    st = p.structure
    cover = {k: np.zeros(n, dtype=int) for k, n in (
        ('g', p.ncons), ('jac', p.nnzjac), ('hess', p.nnzhess),
        ('grad', p.ndec))}
    for rank in range(world):
        sh = sharding.TimeShard(st, p.N, rank, world)
        assert sh.glob.ndec == p.ndec and sh.glob.ncons == p.ncons
        assert sh.glob.nnz_jac == p.nnzjac and sh.glob.nnz_hess == p.nnzhess
        for kind_, c in cover.items():
            for g0, l0, n in sh._pairs(kind_):
                c[g0:g0 + n] += 1
        # local x rows are the global rows k0 .. k1 (+ halo)
        ld = sh.local_dvec(dvec)
        xi = st.var_names.index('x')
        rows = sh.loc.var_rows[xi]
        assert rows == sh.n_local + sh.halo
        np.testing.assert_array_equal(
            ld[sh.loc.var_off[xi]:sh.loc.var_off[xi] + rows * 2].reshape(
                rows, 2),
            p.variables(dvec)['x'][sh.k0:sh.k0 + rows])
    for kind_, c in cover.items():
        assert (c == 1).all(), kind_


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, kind, dims, N, out):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        exp, p, dvec, lam, sigma = _setup(kind, dims, N)
        st = p.structure
        sh = sharding.TimeShard(st, p.N, rank, world)
        y, u = sh.local_data()[:2] if st.data[0]['name'] == 'y' \
            else sh.local_data()[1::-1]
        o = ref_models.make_problem(kind, y, u, dims[0], dt=0.05,
                                    halo=sh.halo)
        assert (o.ndec, o.ncons) == (sh.loc.ndec, sh.loc.ncons)
        ld = sh.local_dvec(dvec)
        ll = sh.local_multipliers(lam)
        grad = o.obj_grad(ld)
        # the all-reduced block: objective + parameter entries of the gradient
        pidx = np.concatenate([
            sh.loc.var_off[i] + np.arange(v['core'])
            for i, v in enumerate(st.vars) if not v['per_sample']])
        red = torch.from_numpy(np.concatenate([[o.obj(ld)], grad[pidx]]))
        dist.all_reduce(red)
        red = red.numpy()
        grad[pidx] = red[1:]
        res = {'grad': grad, 'g': o.constr(ld), 'jac': o.constr_jac_val(ld),
               'hess': o.lag_hess_val(ld, sigma, ll)}
        parts = {k: sh.scatter(k, v, np.zeros(sh.global_size(k)))
                 for k, v in res.items()}
        gathered = [None] * world
        dist.gather_object(parts, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            full = {k: sum(g[k] for g in gathered) for k in parts}
            full['f'] = float(red[0])
            ref = ref_models.make_problem(kind, exp['y'], exp['u'], dims[0],
                                          dt=0.05)
            np.testing.assert_allclose(full['f'], ref.obj(dvec), rtol=1e-13)
            np.testing.assert_allclose(full['grad'], ref.obj_grad(dvec),
                                       rtol=1e-13, atol=1e-300)
            np.testing.assert_allclose(full['g'], ref.constr(dvec),
                                       rtol=1e-13, atol=1e-13)
            np.testing.assert_array_equal(full['jac'],
                                          ref.constr_jac_val(dvec))
            np.testing.assert_array_equal(full['hess'],
                                          ref.lag_hess_val(dvec, sigma, lam))
            out.put('ok')
    except Exception as exc:        # pragma: no cover
        out.put(f'rank {rank}: {exc!r}')
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('kind,dims', [('innovation', (2, 1, 2)),
                                       ('ml_balanced', (2, 1, 2))])
def test_two_rank_gloo_matches_unsharded(kind, dims):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main,
                         args=(r, 2, port, kind, dims, 41, out))
             for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs)
    assert out.get(timeout=5) == 'ok'
