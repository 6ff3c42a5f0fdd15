#!/usr/bin/env python3
"""Host-ingest saturation: N processes (one per GPU) each copy a 520 MB device
buffer -- the D2H bytes of one callback set at N = 1e6 -- to page-locked host
memory at the same time.  Separates the platform's aggregate device->host
ceiling from the design of sharding.SharedVectors:

  private   every rank copies into its own cudaHostAlloc block;
  shared    every rank copies into ITS slice of one shared-memory segment that
            all ranks map and page-lock with cudaHostRegister (what
            SolverFacingEvaluator does).

    python -m torch.distributed.run --nproc-per-node N tools/d2h_saturation.py

One JSON line per mode (rank 0): per-rank and aggregate GB/s.
"""
import ctypes
import json
import mmap
import os
import sys

import torch
import torch.distributed as dist

BYTES = 520_003_472 // 8 * 8
REPS = 10


def timed_copies(dev, host, stream):
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(2):
            host.copy_(dev, non_blocking=True)
        stream.synchronize()
        dist.barrier()
        ev0.record()
        for _ in range(REPS):
            host.copy_(dev, non_blocking=True)
        ev1.record()
        ev1.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3


def main():
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    torch.cuda.set_device(local)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = BYTES // 8
    dev = torch.ones(n, dtype=torch.float64, device=f'cuda:{local}')
    stream = torch.cuda.Stream()
    rt = torch.cuda.cudart()
    for mode in ('private', 'shared'):
        if mode == 'private':
            host = torch.empty(n, dtype=torch.float64).pin_memory()
        else:
            box = [None]
            if rank == 0:
                path = f'/dev/shm/cfem_sat_{os.getpid()}'
                fd = os.open(path, os.O_CREAT | os.O_RDWR, 0o600)
                os.ftruncate(fd, BYTES * world)
                box[0] = path
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                fd = os.open(box[0], os.O_RDWR)
            seg = mmap.mmap(fd, BYTES * world)
            os.close(fd)
            whole = torch.frombuffer(seg, dtype=torch.float64)
            whole[rank * n:(rank + 1) * n] = 0.0        # first touch: own slice
            dist.barrier()
            if rank == 0:
                os.unlink(box[0])
            base = ctypes.addressof(ctypes.c_char.from_buffer(seg))
            err = rt.cudaHostRegister(base, BYTES * world, 0)
            assert int(err) == 0, f'cudaHostRegister failed: {err}'
            host = whole[rank * n:(rank + 1) * n]
        secs = timed_copies(dev, host, stream)
        gbs = BYTES * REPS / secs / 1e9
        all_gbs = [None] * world
        all_secs = [None] * world
        dist.all_gather_object(all_gbs, gbs)
        dist.all_gather_object(all_secs, secs)
        if rank == 0:
            print(json.dumps({
                'mode': mode, 'n_gpus': world, 'bytes_per_copy': BYTES,
                'reps': REPS, 'per_rank_gbs': [round(g, 2) for g in all_gbs],
                'aggregate_gbs': BYTES * REPS * world / max(all_secs) / 1e9,
                'host_cores': os.cpu_count()}), flush=True)
        if mode == 'shared':
            rt.cudaHostUnregister(base)
            del host, whole
        dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
