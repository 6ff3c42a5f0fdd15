"""One solver process in front of a time-sharded problem
(sharding.SolverFacingEvaluator), world_size 2 over gloo on the CPU.

The product has no CPU evaluator: here every rank's ``handle`` is a stand-in
that evaluates its shard with the CPU oracle but moves data exactly like
``backend.Handle`` does (``upload_pieces`` / ``fetch_pieces`` between the
shared global-order vectors and rank-local arrays).  What is tested is the
host logic: the piece maps, the shared-memory request protocol, the IPOPT
callback grouping and a complete interior-point solve through it."""

import os
import socket

import numpy as np
import pytest

from colloc_fem_code_b200 import families, nlp, sharding, synthetic
from oracle import ref_models


class OracleShardHandle:
    """``backend.Handle`` look-alike on the CPU oracle (tests only)."""

    def __init__(self, oracle_problem, shard, allreduce):
        self.o, self.sh, self.allreduce = oracle_problem, shard, allreduce
        self.x = np.full(shard.loc.ndec, np.nan)
        self.lam = np.full(shard.loc.ncons, np.nan)
        self.sigma = 1.0
        self.res = {}
        st = shard.st
        self.pidx = np.concatenate([
            shard.loc.var_off[i] + np.arange(v['core'])
            for i, v in enumerate(st.vars) if not v['per_sample']])
        self.evals = 0

    def upload_pieces(self, which, host_vector, pieces):
        dst = self.x if which == 32 else self.lam
        for l0, g0, n in pieces:
            dst[l0:l0 + n] = host_vector[g0:g0 + n]

    def set_obj_factor(self, sigma):
        self.sigma = float(sigma)

    def eval(self, mask):
        o, x = self.o, self.x
        assert not np.isnan(x).any()
        self.evals += 1
        res = self.res = {}
        if mask & 3:
            grad = o.obj_grad(x)
            red = np.concatenate([[o.obj(x)], grad[self.pidx]])
            self.allreduce(red)
            grad[self.pidx] = red[1:]
            res[1], res[2] = np.array([red[0]]), grad
        if mask & 4:
            res[4] = o.constr(x)
        if mask & 8:
            res[8] = o.constr_jac_val(x)
        if mask & 16:
            assert not np.isnan(self.lam).any()
            res[16] = o.lag_hess_val(x, self.sigma, self.lam)

    def fetch_pieces(self, which, host_vector, pieces):
        src = self.res[which]
        for l0, g0, n in pieces:
            host_vector[g0:g0 + n] = src[l0:l0 + n]

    def synchronize(self):
        pass


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make(kind, N, nx, nu, ny, seed=3):
    exp = synthetic.experiment(seed, N, nx, nu, ny)
    p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
    return exp, p


def _evaluator(p, kind, nx, rank, world):
    import torch
    import torch.distributed as dist
    st = p.structure
    sh = sharding.TimeShard(st, p.N, rank, world)
    data = dict(zip([d['name'] for d in st.data], sh.local_data()))
    o = ref_models.make_problem(kind, data['y'], data['u'], nx, dt=0.05,
                                halo=sh.halo)

    def allreduce(vec):
        t = torch.from_numpy(vec)
        dist.all_reduce(t)

    h = OracleShardHandle(o, sh, allreduce)
    ev = sharding.SolverFacingEvaluator(
        p, sh, h, rank, world,
        broadcast=lambda box: dist.broadcast_object_list(box, src=0),
        barrier=dist.barrier)
    return ev, h


def _rank_main(rank, world, port, case, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        if case == 'callbacks':
            kind, (nx, nu, ny), N = 'ml_balanced', (2, 1, 2), 41
            exp, p = _make(kind, N, nx, nu, ny)
            ev, h = _evaluator(p, kind, nx, rank, world)
            if rank != 0:
                ev.serve()
                assert h.evals == 6
            else:
                assert not os.path.exists(str(ev.sv.path))  # name is gone
                dvec, lam, sigma = synthetic.evaluation_point(p, exp, 3)
                ref = ref_models.make_problem(kind, exp['y'], exp['u'], nx,
                                              dt=0.05)
                f, g = ev.eval_fg(dvec)
                np.testing.assert_allclose(f, ref.obj(dvec), rtol=1e-13)
                np.testing.assert_allclose(g, ref.constr(dvec), rtol=1e-13,
                                           atol=1e-13)
                d2 = dvec * (1 + 1e-3)
                f, grad, g, jv, hv = ev.eval_all(d2, 0.7, lam)
                np.testing.assert_allclose(f, ref.obj(d2), rtol=1e-13)
                np.testing.assert_allclose(grad, ref.obj_grad(d2),
                                           rtol=1e-13, atol=1e-300)
                np.testing.assert_allclose(g, ref.constr(d2), rtol=1e-13,
                                           atol=1e-13)
                np.testing.assert_array_equal(jv, ref.constr_jac_val(d2))
                np.testing.assert_array_equal(
                    hv, ref.lag_hess_val(d2, 0.7, lam))
                # IPOPT's one-at-a-time callbacks: 3 kernel groups per point
                d3 = dvec * (1 - 2e-3)
                fo = np.zeros(1)
                go = np.zeros(p.ncons)
                ev.ipopt_eval(1, d3, True, fo)
                ev.ipopt_eval(4, d3, False, go)
                grad = np.zeros(p.ndec)
                jv = np.zeros(p.nnzjac)
                hv = np.zeros(p.nnzhess)
                ev.ipopt_eval(2, d3, False, grad)
                ev.ipopt_eval(8, d3, False, jv)
                ev.ipopt_eval(16, d3, False, hv, sigma=-1.0, lam=2 * lam)
                assert ev.kernel_groups == 3
                np.testing.assert_allclose(fo[0], ref.obj(d3), rtol=1e-13)
                np.testing.assert_allclose(go, ref.constr(d3), rtol=1e-13,
                                           atol=1e-13)
                np.testing.assert_allclose(grad, ref.obj_grad(d3),
                                           rtol=1e-13, atol=1e-300)
                np.testing.assert_array_equal(jv, ref.constr_jac_val(d3))
                np.testing.assert_array_equal(
                    hv, ref.lag_hess_val(d3, -1.0, 2 * lam))
                # a new multiplier vector at the same x: one more group
                ev.ipopt_eval(16, d3, False, hv, sigma=1.0, lam=lam)
                np.testing.assert_array_equal(
                    hv, ref.lag_hess_val(d3, 1.0, lam))
                assert ev.kernel_groups == 4 and h.evals == 6
                ev.close()
                out.put('ok')
        else:
            from nlp_helpers import (OracleEvaluator, attas_like_experiment,
                                     innovation_setup)
            from colloc_fem_code_b200 import models
            exp = attas_like_experiment(0, 60)
            p = families.make_problem('innovation', exp['y'], exp['u'], 2,
                                      dt=0.05)
            ev, h = _evaluator(p, 'innovation', 2, rank, world)
            if rank != 0:
                ev.serve()
            else:
                dec0, db, cb, scal = innovation_setup(
                    p, exp, models.tril_diag)
                s = nlp.InteriorPointSolver(ev, db, cb)
                s.add_num_option('tol', 1e-9)
                s.set_scaling(*scal)
                x_sh, info_sh = s.solve(dec0)
                ev.close()
                o = ref_models.make_problem('innovation', exp['y'], exp['u'],
                                            2, dt=0.05)
                s2 = nlp.InteriorPointSolver(OracleEvaluator(o), db, cb)
                s2.add_num_option('tol', 1e-9)
                s2.set_scaling(*scal)
                x_1, info_1 = s2.solve(dec0)
                assert info_sh['status'] == info_1['status'] == 'solved'
                assert info_sh['iterations'] == info_1['iterations']
                np.testing.assert_allclose(x_sh, x_1, rtol=1e-8, atol=1e-10)
                out.put('ok')
    except Exception as exc:        # pragma: no cover
        out.put(f'rank {rank}: {exc!r}')
        if rank == 0:
            try:
                ev.stop()
            except Exception:
                pass
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', ['callbacks', 'solve'])
def test_one_solver_two_ranks_gloo(case):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, case, out))
             for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    for pr in procs:
        if pr.is_alive():
            pr.kill()
    assert out.get(timeout=5) == 'ok'
    assert all(pr.exitcode == 0 for pr in procs)


def test_pieces_cover_inputs_and_results():
    exp, p = _make('ml', 23, 2, 1, 2)
    st = p.structure
    for world in (1, 2, 3):
        seen = {k: np.zeros(n, dtype=int) for k, n in (
            ('f', 1), ('grad', p.ndec), ('g', p.ncons), ('jac', p.nnzjac),
            ('hess', p.nnzhess))}
        for rank in range(world):
            sh = sharding.TimeShard(st, p.N, rank, world)
            for kind, size in (('dvec', sh.loc.ndec), ('lam', sh.loc.ncons)):
                local = np.zeros(size, dtype=int)
                for l0, g0, n in sh.input_pieces(kind):
                    local[l0:l0 + n] += 1
                assert (local == 1).all()       # every local entry is fed
            for kind, c in seen.items():
                for l0, g0, n in sh.result_pieces(kind):
                    c[g0:g0 + n] += 1
        for kind, c in seen.items():
            assert (c == 1).all(), (kind, world)


def test_shared_vectors_backing_stores(monkeypatch):
    """tmpfs file when /dev/shm has room, anonymous memfd (opened by the other
    ranks through /proc/<pid>/fd) when it has not; the name disappears after
    unlink() while the mappings stay valid."""
    sizes = {k: 1000 for k in sharding.SharedVectors.FIELDS}
    a = sharding.SharedVectors(sizes, 2)
    assert a.path.startswith('/dev/shm/')
    b = sharding.SharedVectors(sizes, 2, path=a.path)
    a.jac[:] = 3.0
    a.sigma = -1.5
    assert (b.jac == 3.0).all() and b.sigma == -1.5
    for name in sharding.SharedVectors.FIELDS:       # 64-byte aligned fields
        assert getattr(a, name).ctypes.data % 64 == a.ctrl.ctypes.data % 64
    path = a.path
    a.unlink()
    assert not os.path.exists(path)
    b.hess[-1] = 7.0
    assert a.hess[-1] == 7.0
    a.close()
    b.close()

    class Full:
        f_bavail, f_frsize = 1, 4096
    monkeypatch.setattr(os, 'statvfs', lambda p: Full())
    c = sharding.SharedVectors(sizes, 2)
    assert c.path.startswith('/proc/')
    d = sharding.SharedVectors(sizes, 2, path=c.path)
    c.dvec[:] = 2.0
    assert (d.dvec == 2.0).all()
    c.unlink()
    d.grad[0] = 1.0
    assert c.grad[0] == 1.0
    c.close()
    d.close()
