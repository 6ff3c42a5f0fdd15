"""Multi-GPU runs of the time-sharded evaluation at world = 2, 4 and 8 (each
skipped with fewer GPUs): results assembled from all ranks -- with the fused
in-kernel peer-memory reduction (synchronous and pipelined) and with NCCL
all_reduce -- equal the single-GPU evaluation, and every rank holds the same
bits of the reduced objective / parameter gradient."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, mode, out):
    if mode == 'peer_pipelined_fused':      # exchange inside the per-sample
        os.environ['CFEM_SIDE_EXCHANGE'] = '0'      # kernel, not beside the next
    many_tiles = mode.endswith('_many_tiles')
    if many_tiles:
        # several tiles per (persistent) CTA: one resident set of CTAs only
        os.environ['CFEM_WAVES'] = '1'
        mode = mode[:-len('_many_tiles')]
    import torch
    import torch.distributed as dist
    from colloc_fem_code_b200 import backend, families, sharding, synthetic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        nx, nu, ny, N = 2, 1, 2, 20001
        if many_tiles:
            N = 200_001 * world         # > 1 184 tiles per rank
        exp = synthetic.experiment(5, N, nx, nu, ny)
        p = families.make_problem('ml', exp['y'], exp['u'], nx)
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        ev = sharding.ShardedEvaluator(p, rank, world, device=rank)
        h = ev.handle
        stream = torch.cuda.Stream(device=rank)
        torch.cuda.set_stream(stream)
        h.set_stream(stream.cuda_stream)
        if mode.startswith('peer'):
            ev.enable_peer_reduce(pipelined=mode.startswith('peer_pipelined'))
        res = {}
        # several epochs of the hand-shake; the pipelined exchange runs ahead
        # of its collects (ring of 4 epochs) and is only joined by a fetch
        pipelined = mode.startswith('peer_pipelined')
        reps = 11 if pipelined else 3
        for rep in range(reps):
            ev.set_point(dvec + 1e-3 * rep, sigma, lam)
            h.eval(backend.ALL)
            if pipelined and rep % 4 != 3 and rep != reps - 1:
                continue
            if mode == 'nccl':
                ptr = h.device_ptrs()['reduce']

                class Arr:
                    __cuda_array_interface__ = {
                        'shape': (ev.n_reduce,), 'typestr': '<f8',
                        'data': (ptr, False), 'version': 2}
                red = torch.as_tensor(Arr(), device=f'cuda:{rank}')
                dist.all_reduce(red)
                h.apply_reduced(ptr)
            res = {k: h.fetch(b) for k, b in (('f', 1), ('grad', 2), ('g', 4),
                                              ('jac', 8), ('hess', 16))}
        parts = {k: ev.shard.scatter(k, res[k],
                                     np.zeros(ev.shard.global_size(k)))
                 for k in ('grad', 'g', 'jac', 'hess')}
        parts['f'] = float(res['f'][0])
        # the reduced parameter block of the gradient, as this rank holds it
        slots = ev.lib.model['reduce'][1:]
        st = p.structure
        loc_off = dict(zip(st.var_names, h.layout()['var_offset']))
        parts['red'] = np.array([res['grad'][loc_off[st.var_names[v]] + fl]
                                 for v, fl in slots])
        gathered = [None] * world
        dist.gather_object(parts, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            d = dvec + 1e-3 * (reps - 1)
            ref = {'f': p.obj(d), 'grad': p.obj_grad(d), 'g': p.constr(d),
                   'jac': p.constr_jac_val(d),
                   'hess': p.lag_hess_val(d, sigma, lam)}
            for other in gathered[1:]:                      # same bits
                assert gathered[0]['f'] == other['f']
                np.testing.assert_array_equal(other['red'],
                                              gathered[0]['red'])
            np.testing.assert_allclose(gathered[0]['f'], ref['f'], rtol=1e-13)
            for k in ('g', 'jac', 'hess'):
                full = sum(g[k] for g in gathered)
                np.testing.assert_array_equal(full, ref[k], err_msg=k)
            full = sum(g['grad'] for g in gathered)
            np.testing.assert_allclose(full, ref['grad'], rtol=1e-13,
                                       atol=1e-300)
            out.put('ok')
    except Exception as exc:            # pragma: no cover
        out.put(f'rank {rank}: {exc!r}')
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('mode', ['peer', 'peer_pipelined',
                                  'peer_pipelined_fused', 'nccl',
                                  'peer_many_tiles',
                                  'peer_pipelined_many_tiles'])
def test_multi_gpu_sharded_equals_single(mode, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, mode, out))
             for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(600)
    msg = out.get(timeout=10)
    assert msg == 'ok', msg
    assert all(pr.exitcode == 0 for pr in procs)


def _solver_rank_main(rank, world, port, mode, out):
    """Rank 0 drives the callbacks (and a whole interior-point solve) through
    sharding.SolverFacingEvaluator; every rank moves only its own pieces."""
    import torch
    import torch.distributed as dist
    from colloc_fem_code_b200 import families, nlp, sharding, synthetic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    ev = None
    try:
        nx, nu, ny, N = 2, 1, 2, 30001
        exp = synthetic.experiment(7, N, nx, nu, ny)
        p = families.make_problem('ml', exp['y'], exp['u'], nx)
        ev = sharding.solver_facing_evaluator(p, rank, world, device=rank,
                                              reduce=mode)
        if rank != 0:
            ev.serve()
        else:
            assert ev.pinned
            from oracle import ref_models
            ref = ref_models.make_problem('ml', exp['y'], exp['u'], nx)
            dvec, lam, sigma = synthetic.evaluation_point(p, exp)
            for rep in range(3):
                d = dvec * (1 + 1e-3 * rep)
                f, g = ev.eval_fg(d)
                np.testing.assert_allclose(f, ref.obj(d), rtol=1e-12)
                f, grad, g, jv, hv = ev.eval_all(d, sigma, lam)
                np.testing.assert_allclose(f, ref.obj(d), rtol=1e-12)
                np.testing.assert_allclose(grad, ref.obj_grad(d), rtol=1e-12,
                                           atol=1e-300)
                scale = 1e-12 * (1 + np.abs(d).max())
                np.testing.assert_allclose(g, ref.constr(d), rtol=1e-12,
                                           atol=scale)
                np.testing.assert_allclose(jv, ref.constr_jac_val(d),
                                           rtol=1e-12, atol=1e-300)
                np.testing.assert_allclose(hv, ref.lag_hess_val(d, sigma, lam),
                                           rtol=1e-12, atol=1e-300)
            hv2 = np.zeros(p.nnzhess)
            ev.ipopt_eval(16, d, False, hv2, sigma=-0.5, lam=3 * lam)
            np.testing.assert_allclose(hv2, ref.lag_hess_val(d, -0.5, 3 * lam),
                                       rtol=1e-12, atol=1e-300)
            ev.close()
            out.put('ok')
    except Exception as exc:            # pragma: no cover
        out.put(f'rank {rank}: {exc!r}')
        if rank == 0 and ev is not None:
            try:
                ev.stop()
            except Exception:
                pass
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode,world', [('peer', 2), ('nccl', 2), ('peer', 4),
                                        ('peer', 8)])
def test_multi_gpu_one_solver_process(mode, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solver_rank_main,
                         args=(r, world, port, mode, out))
             for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(600)
    for pr in procs:
        if pr.is_alive():
            pr.kill()
    msg = out.get(timeout=10)
    assert msg == 'ok', msg
    assert all(pr.exitcode == 0 for pr in procs)


def test_one_gpu_solver_facing_pieces():
    """world = 1: the piecewise transfers (cfem_upload_pieces /
    cfem_fetch_pieces) against the whole-vector path of the same handle."""
    from colloc_fem_code_b200 import backend, families, sharding, synthetic
    nx, nu, ny, N = 5, 3, 3, 513
    exp = synthetic.experiment(11, N, nx, nu, ny)
    p = families.make_problem('ml_balanced', exp['y'], exp['u'], nx)
    dvec, lam, sigma = synthetic.evaluation_point(p, exp)
    sev = sharding.ShardedEvaluator(p, 0, 1)
    ev = sharding.SolverFacingEvaluator(
        p, sev.shard, sev.handle, 0, 1, broadcast=lambda box: None,
        barrier=lambda: None, lib=sev.lib)
    assert ev.pinned
    f, grad, g, jv, hv = ev.eval_all(dvec, sigma, lam)
    h = sev.handle
    h.set_dvec(dvec)
    h.set_multipliers(sigma, lam)
    h.eval(backend.ALL)
    assert f == h.fetch(backend.F)[0]
    for bit, got in ((backend.GRAD, grad), (backend.G, g), (backend.JAC, jv),
                     (backend.HESS, hv)):
        np.testing.assert_array_equal(got, h.fetch(bit))
    f2, g2 = ev.eval_fg(dvec * 1.01)
    h.set_dvec(dvec * 1.01)
    h.eval(backend.F | backend.G)
    assert f2 == h.fetch(backend.F)[0]
    np.testing.assert_array_equal(g2, h.fetch(backend.G))
    ev.close()
