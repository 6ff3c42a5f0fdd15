#!/usr/bin/env python3
"""Device-resident throughput of one full callback set for every BASELINE
configuration shape, as a JSON-lines table:
kernel time (CUDA events, L2 flushed), algorithmic GB/s, fraction of the
measured HBM peak, whole-step callback sets/s and samples/s."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from colloc_fem_code_b200 import backend, families, synthetic  # noqa: E402

CONFIGS = [
    ('C1 hfb320-equivalent', 'ndisc_zoh', (4, 2, 7), 601),
    ('C1 hfb320-equivalent', 'ndisc_zoh', (4, 2, 7), 1_000_000),
    ('C2 attas_sp_ml', 'ml', (2, 1, 2), 1000),
    ('C2 attas_sp_ml', 'ml', (2, 1, 2), 1_000_000),
    ('C2 attas_sp_innov_bal', 'balanced', (2, 1, 2), 1_000_000),
    ('C3 blackbox_innov_bal', 'balanced', (5, 3, 3), 250),
    ('C3 blackbox_innov_bal', 'balanced', (5, 3, 3), 1_000_000),
    ('C5 long trajectory', 'ml', (2, 1, 2), 10_000_000),
    ('extension trapezoid', 'trapezoid', (2, 1, 2), 1_000_000),
]


def structure_bytes_per_sample(st):
    """Compulsory traffic per sample from the structure tables: every
    per-sample input once (variables, data, multipliers), every per-sample
    output once (constraint values, gradient block, Jacobian and Hessian
    values)."""
    n = sum(v['core'] for v in st.vars if v['per_sample'])
    n += sum(d['core'] for d in st.data)
    for f in st.funs:
        if not f['per_sample']:
            continue
        if f['is_objective']:
            n += sum(st.vars[r[1]]['core'] for a, r in f['args'].items()
                     if r[0] == 'var' and a in f['spec'].jac)   # gradient block
        else:
            n += 2 * f['out_core']          # multipliers in, values out
    n += sum(b['c'] for b in st.jac_blocks if st.funs[b['fun']]['per_sample'])
    n += sum(b['c'] for b in st.hess_blocks if st.funs[b['fun']]['per_sample'])
    return 8 * n


def main():
    peak, _ = bench.measured_peak()
    for label, kind, dims, N in CONFIGS:
        nx, nu, ny = dims
        exp = synthetic.experiment(0, N, nx, nu, ny)
        p = families.make_problem(kind, exp['y'], exp['u'], nx, dt=0.05)
        dvec, lam, sigma = synthetic.evaluation_point(p, exp)
        del exp
        h = p.backend.handle
        h.set_kernel_timing(True)
        h.set_dvec(dvec)
        h.set_multipliers(sigma, lam)
        dptr = h.device_ptrs()['dvec']
        steps = 20
        for i in range(5):
            h.flush_l2(256 << 20)
            h.set_dvec_device(dptr)
            h.eval(31)
        h.synchronize()
        for i in range(steps):      # pass 1: events around the kernel alone
            h.flush_l2(256 << 20)
            h.set_dvec_device(dptr)
            h.eval(31)
        kms = float(np.median(h.sample_kernel_ms_history(steps)))
        h.set_kernel_timing(False)
        step_ms = []
        for i in range(steps + 3):  # pass 2: the whole step, no inner events
            h.flush_l2(256 << 20)
            h.event_record(0)
            h.set_dvec_device(dptr)
            h.eval(31)
            h.event_record(1)
            step_ms.append(h.event_elapsed_ms(0, 1))
        step_ms = step_ms[3:]
        balg = structure_bytes_per_sample(p.structure)
        if kind != 'trapezoid':     # the closed form of SURVEY.md section 8(d)
            assert balg == bench.algorithmic_bytes_per_sample(nx, nu, ny)
        gbs = balg * N / (kms * 1e-3) / 1e9
        sms = float(np.median(step_ms))
        extra = {}
        if N <= 10_000:     # native script lengths: host-API latency + CPU
            import time
            from oracle import ref_models
            host = backend.HostBuffers(h)
            host.dvec[:] = dvec
            host.lam[:] = lam
            ts = []
            for i in range(30):
                t0 = time.perf_counter()
                host.upload(sigma)
                h.eval(31)
                host.fetch_all()
                ts.append(time.perf_counter() - t0)
            extra['e2e_host_api_ms'] = 1e3 * float(np.median(ts[5:]))
            o = ref_models.make_problem(kind, p.y, p.u, nx, dt=0.05)
            tc = []
            for i in range(4):
                t0 = time.perf_counter()
                o.obj(dvec), o.obj_grad(dvec), o.constr(dvec)
                o.constr_jac_val(dvec), o.lag_hess_val(dvec, sigma, lam)
                tc.append(time.perf_counter() - t0)
            extra['cpu_oracle_ms'] = 1e3 * min(tc[1:])
        print(json.dumps({
            **extra, 'config': label, 'family': kind, 'dims': dims, 'N': N,
            'kernel_ms': kms, 'step_ms': sms, 'algorithmic_GBps': gbs,
            'frac_of_measured_peak': gbs / peak,
            'callback_sets_per_s': 1e3 / sms,
            'samples_per_s': N * 1e3 / sms,
            'algorithmic_bytes_per_sample': balg}), flush=True)
        p.backend.close()


if __name__ == '__main__':
    main()
