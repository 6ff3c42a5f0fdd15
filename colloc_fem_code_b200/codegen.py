"""CUDA code generator: problem structure -> one fused FP64 translation unit.

This replaces the reference's sympy -> NumPy code generation
(``sym2num`` / ``symoptim.compile_class``; call sites
/root/reference/attas_sp_ml.py:85-86, /root/reference/mc_blackbox_cfem.py:86-95)
with a generator of sm_100a CUDA.  Input is ``optim.Structure`` (N-independent);
output is the source of ONE shared library that exports the C ABI of
``include/cfem.h``.  The generated text contains only what depends on the
model: the structure tables and the straight-line expression code of every
structural nonzero.  The machinery around it (staging, per-warp output
transposition, reductions, the whole host side) is the hand-written code in
``csrc/cfem_device.cuh`` and ``csrc/cfem_host.inl``.

Kernels per library
  cfem_sample_kernel_m<mask>  one thread per sample, one CTA per tile of
      CFEM_TILE samples and (blockIdx.y) per problem of a batch; evaluates every
      per-sample function selected by ``mask`` (objective partial sums,
      gradient block, constraint values, Jacobian blocks, Hessian blocks) from
      one staging of the tile's inputs;
  cfem_param_kernel           one thread per entry of the parameter-only
      functions (values, Jacobian, Hessian);
  (finalisation)              the last CTA of a problem to retire does the
      fixed-order reduction over tiles + the sample-independent objective
      terms (times the row count) inside the per-sample kernel itself;
  cfem_apply_reduced_kernel   writes an externally all-reduced objective /
      parameter-gradient vector back (time-sharded multi-GPU runs).
"""

import json
import os

from . import symoptim

F, GRAD, G, JAC, HESS = 1, 2, 4, 8, 16
ALL = 31

#: kernel variants that are instantiated (a request is served by the smallest
#: superset): single callbacks, IPOPT's usual groupings and the full set.
DEFAULT_MASKS = (F, F | GRAD, G, JAC, HESS, F | G, F | GRAD | G | JAC, ALL)
# defaults of the two round-2 kernel options (Generator.__init__); the
# alternatives stay selectable (CFEM_RETIRE / CFEM_EARLY_LOADS at build time)
DEFAULT_RETIRE = 'tree'
DEFAULT_EARLY_LOADS = 0


def _skew_size(c, rows):
    """Doubles of a [rows][c] array in the Skew<c> layout of cfem_device.cuh,
    rounded up to an even count (16-byte alignment of what follows)."""
    import math
    m = 16 // math.gcd(c, 16)
    n = rows * c + (rows + m - 1) // m
    return n + (n & 1)


class _Item:
    """One output block a sample kernel produces for one function."""

    def __init__(self, c, codes, dest, mult=None, uniform=False, deps=()):
        self.c = c              # doubles per sample
        self.codes = codes      # C expression per element
        self.dest = dest        # C expression: pointer to row 0 of the block
        self.mult = mult        # optional C expressions multiplied in
        # every entry is the same for all samples (constants / parameters /
        # obj_factor): the block is a periodic pattern, streamed from a small
        # per-CTA table without the per-sample transposition
        self.uniform = uniform
        self.deps = list(deps)  # (arg, flat) symbols the entries use
        self.pat_off = None     # offset of the pattern in the CTA's table


class Generator:
    def __init__(self, structure, tile=None, pass_budget=None, masks=None,
                 min_blocks=None, experiment=None, store=None, retire=None,
                 early_loads=None):
        self.st = structure
        self.masks = tuple(masks or DEFAULT_MASKS)
        self.funs = structure.funs
        self.sample_funs = [i for i, f in enumerate(self.funs)
                            if f['per_sample']]
        self.param_funs = [i for i, f in enumerate(self.funs)
                           if not f['per_sample']]
        # defaults from the B200 sweeps (profiles/README.md): 128-sample tiles,
        # output passes of <= 6 doubles per lane
        if tile is None:
            tile = int(os.environ.get('CFEM_TILE', 0)) or 128
        if pass_budget is None:
            pass_budget = int(os.environ.get('CFEM_PASS_BUDGET', 0)) or 6
        assert tile % 32 == 0 and 32 <= tile <= 1024
        self.tile = tile
        self.pass_budget = pass_budget
        self.min_blocks = min_blocks or int(
            os.environ.get('CFEM_MIN_BLOCKS', 0))
        # tuning experiments only (tools/sweep.py): 'nostore' / 'noload'
        self.experiment = experiment
        self.store = store          # None (st.global.cs) | 'wb' | 'cg' | 'wt'
        # retirement of the tile CTAs' partial sums: 'tree' = "last block
        # done" tree (fence + ticket per CTA), 'flag' = self-validating slots
        # polled by a finaliser CTA (cfem_device.cuh, "fence-free retirement")
        self.retire = retire or os.environ.get('CFEM_RETIRE') or DEFAULT_RETIRE
        assert self.retire in ('tree', 'flag')
        # issue the first tile's cp.async loads BEFORE the parameters are
        # staged and the pattern table is built (one prologue barrier)
        if early_loads is None:
            early_loads = int(os.environ.get('CFEM_EARLY_LOADS',
                                             DEFAULT_EARLY_LOADS))
        self.early_loads = bool(early_loads)
        self._reduce_slots()

    # ------------------------------------------------------------------
    # classification helpers
    # ------------------------------------------------------------------
    def _is_sample_dep(self, fun, deps):
        return any(fun['args'][a][0] in ('var', 'data') for a, _ in deps)

    def _reduce_slots(self):
        """Slot 0 is the objective; then one slot per structural parameter
        entry of the objective gradient.  ``dyn`` slots have per-sample terms
        and therefore per-tile partial sums."""
        self.slots = [{'var': -1, 'flat': 0, 'dyn': True}]
        self.slot_of = {}
        for fi, f in enumerate(self.funs):
            if not f['is_objective']:
                continue
            spec = f['spec']
            for wrt in spec.wrt:
                ref = f['args'][wrt]
                if ref[0] != 'param' or wrt not in spec.jac:
                    continue
                for e in spec.jac[wrt]:
                    key = (ref[1], e.index[0])
                    dyn = f['per_sample'] and any(
                        self._is_sample_dep(f, d) for _, d in e.terms)
                    if key not in self.slot_of:
                        self.slot_of[key] = len(self.slots)
                        self.slots.append({'var': key[0], 'flat': key[1],
                                           'dyn': False})
                    if dyn:
                        self.slots[self.slot_of[key]]['dyn'] = True
        self.dyn_slots = [i for i, s in enumerate(self.slots) if s['dyn']]
        self.dyn_index = {s: i for i, s in enumerate(self.dyn_slots)}

    # ------------------------------------------------------------------
    # identifiers
    # ------------------------------------------------------------------
    def _define_args(self, fun, deps, where, layout=None):
        """#define / const lines binding v_<arg>_<flat> for ``deps``.

        where: 'sample' (tile in shared memory), 'global' (dvec in HBM).
        Returns (lines, undef_lines).
        """
        lines, undefs = [], []
        for a, flat in sorted(set(deps)):
            ident = symoptim.c_ident(a, flat)
            ref = fun['args'][a]
            if ref[0] == 'param':
                if where == 'sample':
                    off = layout['param_off'][ref[1]] + flat
                    lines.append(f'#define {ident} sp[{off}]')
                else:
                    lines.append(f'#define {ident} dvec[a.var_off[{ref[1]}] '
                                 f'+ {flat}]')
                undefs.append(f'#undef {ident}')
            elif ref[0] == 'scalar':
                lines.append(f'#define {ident} a.scalars[{ref[1]}]')
                undefs.append(f'#undef {ident}')
            else:
                assert where == 'sample'
                key = (ref[0], ref[1])
                st = layout['stor'][key]
                lines.append(
                    f'const double {ident} = {st["name"]}'
                    f'[cfem::Skew<{st["core"]}>::row(row + {ref[2]}) + {flat}];')
        return lines, undefs

    # ------------------------------------------------------------------
    # per-mask plan of a sample kernel
    # ------------------------------------------------------------------
    def _plan(self, mask):
        """Items, reductions and needed symbols per per-sample function."""
        plan = []
        st = self.st
        jac_of = {}
        for bi, b in enumerate(st.jac_blocks):
            jac_of.setdefault(b['fun'], []).append((bi, b))
        hess_of = {}
        for bi, b in enumerate(st.hess_blocks):
            hess_of.setdefault(b['fun'], []).append((bi, b))
        for fi in self.sample_funs:
            f = self.funs[fi]
            spec = f['spec']
            items, reds, deps = [], [], []
            lam_needed = False
            if f['is_objective']:
                if mask & (F | GRAD):
                    for e in spec.values:
                        for code, d in e.terms:
                            if self._is_sample_dep(f, d):
                                reds.append((0, code))
                                deps += d
                    for wrt in spec.wrt:
                        ref = f['args'][wrt]
                        if ref[0] != 'param':
                            continue
                        for e in spec.jac.get(wrt, []):
                            slot = self.slot_of[(ref[1], e.index[0])]
                            for code, d in e.terms:
                                if self._is_sample_dep(f, d):
                                    reds.append((slot, code))
                                    deps += d
                if mask & GRAD:
                    for wrt in spec.wrt:
                        ref = f['args'][wrt]
                        if ref[0] != 'var' or wrt not in spec.jac:
                            continue
                        core = spec.core_size(wrt)
                        codes = ['0.0'] * core
                        for e in spec.jac[wrt]:
                            assert e.index[1] == 0
                            codes[e.index[0]] = e.code
                            deps += e.deps
                        dest = (f'a.grad + b * a.ndec + a.var_off[{ref[1]}] '
                                f'+ {ref[2] * core}')
                        items.append(_Item(core, codes, dest))
            else:
                ci = f['cons_index']
                if mask & G:
                    codes = [e.code for e in spec.values]
                    for e in spec.values:
                        deps += e.deps
                    items.append(_Item(
                        f['out_core'], codes,
                        f'a.g + b * a.ncons + a.cons_off[{ci}]'))
                if mask & JAC:
                    for bi, blk in jac_of.get(fi, []):
                        bdeps = [d for e in blk['entries'] for d in e.deps]
                        uni = blk['c'] > 1 and not self._is_sample_dep(f, bdeps)
                        if not uni:
                            deps += bdeps
                        items.append(_Item(
                            blk['c'], [e.code for e in blk['entries']],
                            f'a.jac + b * a.nnz_jac + a.jac_off[{bi}]',
                            uniform=uni, deps=bdeps))
            if mask & HESS:
                for bi, blk in hess_of.get(fi, []):
                    if f['is_objective']:
                        mult = ['a.obj_factor'] * blk['c']
                    else:
                        mult = [f'lam_{fi}_{e.index[2]}'
                                for e in blk['entries']]
                        lam_needed = True
                    bdeps = [d for e in blk['entries'] for d in e.deps]
                    uni = (f['is_objective'] and blk['c'] > 1
                           and not self._is_sample_dep(f, bdeps))
                    if not uni:
                        deps += bdeps
                    items.append(_Item(
                        blk['c'], [e.code for e in blk['entries']],
                        f'a.hess + b * a.nnz_hess + a.hess_off[{bi}]', mult,
                        uniform=uni, deps=bdeps))
            if items or reds:
                plan.append({'fi': fi, 'items': items, 'reds': reds,
                             'deps': sorted(set(deps)),
                             'lam': lam_needed})
        return plan

    def _smem_layout(self, plan):
        """Shared-memory carve-up (in doubles) of one sample kernel."""
        off = 0
        param_off = {}
        stor = {}
        for p in plan:
            f = self.funs[p['fi']]
            udeps = [d for it in p['items'] if it.uniform for d in it.deps]
            for a, _ in list(p['deps']) + udeps:
                ref = f['args'][a]
                if ref[0] == 'param' and ref[1] not in param_off:
                    param_off[ref[1]] = None
                elif ref[0] in ('var', 'data'):
                    key = (ref[0], ref[1])
                    s = stor.setdefault(key, {'shift': 0})
                    s['shift'] = max(s['shift'], ref[2])
            if p['lam']:
                stor[('lam', p['fi'])] = {'shift': 0}
        for v in sorted(param_off):
            param_off[v] = off
            off += self.st.vars[v]['core']
        off += off & 1
        pat_off = off               # periodic patterns of the uniform blocks
        for p in plan:
            for it in p['items']:
                if it.uniform:
                    it.pat_off = off - pat_off
                    off += it.c
        off += off & 1
        stage_off = off             # start of staging buffer 0
        for key in sorted(stor):
            s = stor[key]
            if key[0] == 'var':
                core = self.st.vars[key[1]]['core']
            elif key[0] == 'data':
                core = self.st.data[key[1]]['core']
            else:
                core = self.funs[key[1]]['out_core']
            s['core'] = core
            s['nrows'] = self.tile + s['shift']
            s['name'] = f's_{key[0]}{key[1]}'
            off += off & 1      # keep 16-byte alignment of every region
            s['off'] = off - stage_off      # relative to the staging buffer
            off += _skew_size(core, s['nrows'])
        off += off & 1
        stage_size = off - stage_off
        # output passes: group the items of every function under the budget
        wbuf = 0
        for p in plan:
            passes, cur, used = [], [], 0
            for it in p['items']:
                if it.uniform:
                    continue        # streamed from the pattern table
                need = 0 if it.c == 1 else it.c + 1
                if cur and used + need > self.pass_budget:
                    passes.append(cur)
                    cur, used = [], 0
                cur.append(it)
                used += need
            if cur:
                passes.append(cur)
            p['passes'] = passes
            p['uniform'] = [it for it in p['items'] if it.uniform]
            for ps in passes:
                wbuf = max(wbuf, sum(0 if it.c == 1 else _skew_size(it.c, 32)
                                     for it in ps))
        warps = self.tile // 32
        off += off & 1
        red_off = off
        off += warps * max(1, len(self.dyn_slots))
        off += off & 1
        wbuf_off = off
        off += warps * wbuf
        # staging buffer 1 (target of the next tile's cp.async) comes LAST: a
        # launch in which every CTA has exactly one tile never touches it and
        # is made with `total_single` bytes of dynamic shared memory only
        off += off & 1
        stage1_off = off
        off += stage_size
        return {'param_off': param_off, 'stor': stor, 'red_off': red_off,
                'wbuf_off': wbuf_off, 'wbuf': wbuf, 'total': off,
                'total_single': stage1_off,
                'stage_off': stage_off, 'stage_size': stage_size,
                'stage_stride': stage1_off - stage_off,
                'pat_off': pat_off}

    # ------------------------------------------------------------------
    # emitters
    # ------------------------------------------------------------------
    def _emit_sample_kernel(self, mask):
        plan = self._plan(mask)
        lay = self._smem_layout(plan)
        T = self.tile
        nred = len(self.dyn_slots)
        w = []
        w.append(f'// mask {mask}: ' + ' '.join(
            n for n, bit in (('F', F), ('GRAD', GRAD), ('G', G), ('JAC', JAC),
                             ('HESS', HESS)) if mask & bit))
        w.append(f'constexpr size_t kSmemBytes_m{mask} = {lay["total"] * 8};')
        w.append(f'constexpr size_t kSmemBytes1_m{mask} = '
                 f'{lay["total_single"] * 8};   // one tile per CTA')
        bounds = f'{T}, {self.min_blocks}' if self.min_blocks else f'{T}'
        w.append(f'__global__ void __launch_bounds__({bounds})')
        w.append(f'cfem_sample_kernel_m{mask}(const __grid_constant__ cfem::KArgs a)')
        w.append('{')
        w.append('    extern __shared__ __align__(16) double smem[];')
        w.append('    const int tid = threadIdx.x, lane = tid & 31, '
                 'warp = tid >> 5;')
        w.append('    const long long b = blockIdx.y;')
        if mask & (F | GRAD):
            w.append('    if (a.finaliser && (long long)blockIdx.x == a.nctas) {')
            w.append('        // no tile: finalise the sums and exchange them '
                     'with the peer GPUs beside the last tiles')
            if self.retire == 'tree':
                w.append('        cfem::wait_groups_done(a, b, tid);')
            w.append(f'        cfem_finalize(a, {mask}u, b, tid, smem + '
                     f'{lay["red_off"]});')
            if self.retire == 'tree':
                w.append('        if (tid == 0) a.done_count[b] = 0u;')
            w.append('        asm volatile("griddepcontrol.wait;" ::: "memory");')
            w.append('        return;')
            w.append('    }')
        w.append('    const double* __restrict__ dvec = a.dvec + b * a.ndec;')
        w.append('    double* const sp = smem;')
        w.append(f'    double* const wb = smem + {lay["wbuf_off"]} + warp * '
                 f'{lay["wbuf"]};')
        w.append('    (void)lane; (void)dvec; (void)sp; (void)wb;')

        def stage(item_expr, buf_expr):
            """cp.async of the rows of work item ``item_expr`` into staging
            buffer ``buf_expr``."""
            out = ['        {',
                   f'            double* const stg = smem + '
                   f'{lay["stage_off"]} + ({buf_expr}) * {lay["stage_stride"]};',
                   '            long long r0; int spw_;',
                   f'            cfem::item_range(a, {item_expr}, r0, spw_);',
                   f'            const int nr = spw_ * {T // 32};',
                   '            (void)stg; (void)nr;']
            for key in sorted(lay['stor']):
                st_ = lay['stor'][key]
                if key[0] == 'var':
                    src = f'dvec + a.var_off[{key[1]}]'
                    rows = f'a.var_rows[{key[1]}]'
                elif key[0] == 'data':
                    src = (f'a.data[{key[1]}] + b * a.data_rows[{key[1]}] * '
                           f'{st_["core"]}')
                    rows = f'a.data_rows[{key[1]}]'
                else:
                    ci = self.funs[key[1]]['cons_index']
                    src = f'a.lam + b * a.ncons + a.cons_off[{ci}]'
                    rows = f'a.fun_rows[{key[1]}]'
                out.append(f'            cfem::stage_rows_async<{st_["core"]}, '
                           f'{st_["nrows"]}>(stg + {st_["off"]}, {src}, '
                           f'{rows}, r0, nr + {st_["shift"]}, tid);')
            out.append('        }')
            return out

        has_red = bool(mask & (F | GRAD))
        # the reduction runs right after the last function that feeds it, in
        # the CTA's LAST tile, i.e. BEFORE the bulk of that tile's stores: its
        # fence then has no store queue to drain and the serial tail of the
        # grid-wide tree (and of the cross-GPU exchange) overlaps the stores
        red_after = max([i for i, p in enumerate(plan) if p['reds']],
                        default=-1) if has_red else -1
        reduce_call = [
            f'        if (cfem::tree_reduce<{max(nred, 1)}>(a, b, red, smem + '
            f'{lay["red_off"]}, tid) && !a.finaliser) {{',
            '            // this CTA retired last: it finalises (fixed '
            'summation tree => deterministic)',
            f'            cfem_finalize(a, {mask}u, b, tid, smem + '
            f'{lay["red_off"]});',
            '            if (tid == 0) a.done_count[b] = 0u;',
            '        }']
        if self.retire == 'flag':
            reduce_call = [
                '        // one store per slot, no fence, no ticket: the finaliser '
                'CTA polls the slots',
                f'        cfem::publish_partial<{max(nred, 1)}>(a, b, red, smem + '
                f'{lay["red_off"]}, tid);']

        # Work items (cfem_args.cuh): CTA c takes items c, c + gridDim.x, ...;
        # the inputs of the next item are in flight (cp.async) while this one
        # is evaluated and streamed out.  Per item: (1) the sample-independent
        # blocks are streamed from the pattern table -- they need no input,
        # so these stores overlap the latency of the item's own loads; (2)
        # wait for the loads; (3) the reduction (last item only); (4) the
        # sample-dependent blocks.
        w.append('    const long long G = a.nctas;      // tile CTAs (the grid may '
                 'hold one more: the finaliser)')
        w.append('    int buf = 0;')
        w.append('    long long item = blockIdx.x;')
        if self.early_loads:
            # the tile's loads depend on nothing but the block index: they go
            # first, the prologue below runs under their latency
            w.append('    if (item < a.nitems)')
            w += stage('item', '0')
            w.append('    cfem::cp_async_commit();')
        for v, off in sorted(lay['param_off'].items()):
            w.append(f'    cfem::stage_contig(sp + {off}, dvec + '
                     f'a.var_off[{v}], {self.st.vars[v]["core"]}, tid);')
        if mask & (F | GRAD):
            w.append(f'    double red[{max(nred, 1)}];')
            w.append(f'    for (int r = 0; r < {max(nred, 1)}; ++r) '
                     'red[r] = 0.0;')
        w.append(f'    double* const pat = smem + {lay["pat_off"]};')
        w.append('    (void)pat;')
        if any(p['uniform'] for p in plan):
            if self.early_loads:
                # the pattern entries read the parameters from global memory
                # (no dependence on the staged copy): one barrier covers both
                w.append(f'    if (tid == {T - 32}) {{   // once per (persistent) '
                         'CTA, beside the staging of warp 0')
            else:
                w.append('    __syncthreads();       // parameters are staged')
                w.append('    if (tid == 0) {        // once per (persistent) CTA')
            for p in plan:
                f = self.funs[p['fi']]
                for it in p['uniform']:
                    defs, undefs = self._define_args(
                        f, it.deps, 'global' if self.early_loads else 'sample',
                        lay)
                    w += defs
                    for j, code in enumerate(it.codes):
                        if it.mult:
                            code = f'({it.mult[j]}) * ({code})'
                        w.append(f'        pat[{it.pat_off + j}] = {code};')
                    w += undefs
            w.append('    }')
            w.append('    __syncthreads();       // parameters staged, pattern '
                     'table built')
        if not self.early_loads:
            w.append('    if (item < a.nitems)')
            w += stage('item', '0')
            w.append('    cfem::cp_async_commit();')
        w.append('    for (; item < a.nitems; item += G, buf ^= 1) {')
        w.append('    const bool last_item = item + G >= a.nitems;')
        w.append('    if (!last_item)')
        w += stage('item + G', 'buf ^ 1')
        w.append('    cfem::cp_async_commit();')
        w.append('    long long k0; int spw;')
        w.append('    cfem::item_range(a, item, k0, spw);')
        w.append('    const long long kw = k0 + warp * spw;    // first sample '
                 'of this warp')
        w.append('    const int row = warp * spw + lane;        // row of this '
                 'thread in the staged item')
        w.append('    (void)kw; (void)row;')
        for p in plan:
            if not p['uniform']:
                continue
            fi = p['fi']
            w.append('    {')
            w.append(f'        const long long left = a.fun_rows[{fi}] - kw;')
            w.append('        const int nvalid = left >= spw ? spw : '
                     '(left > 0 ? (int)left : 0);')
            w.append('        if (nvalid > 0) {')
            for it in p['uniform']:
                w.append(f'            cfem::warp_store_periodic<{it.c}>(pat + '
                         f'{it.pat_off}, lane, ({it.dest}) + kw * {it.c}, '
                         'nvalid);')
            w.append('        }')
            w.append('    }')
        w.append('    cfem::cp_async_wait<1>();      // this item has landed')
        w.append('    __syncthreads();')
        w.append(f'    double* const stg = smem + {lay["stage_off"]} + buf * '
                 f'{lay["stage_stride"]};')
        w.append('    (void)stg;')
        for key in sorted(lay['stor']):
            st_ = lay['stor'][key]
            w.append(f'    const double* const {st_["name"]} = stg + '
                     f'{st_["off"]};')
        for pi, p in enumerate(plan):
            fi = p['fi']
            f = self.funs[fi]
            defs, undefs = self._define_args(f, p['deps'], 'sample', lay)

            def open_block():
                w.append('    {')
                w.append(f'        const long long left = a.fun_rows[{fi}] - kw;')
                w.append('        const int nvalid = left >= spw ? spw : '
                         '(left > 0 ? (int)left : 0);')
                w.append('        const bool act = lane < nvalid;')
                w.append('        (void)act;')
                w.append('        if (nvalid > 0) {')
                w.extend('        ' + d if d.startswith('#')
                         else '            ' + d for d in defs)

            def close_block():
                w.extend('        ' + u for u in undefs)
                w.append('        }')
                w.append('    }')

            w.append(f'    // ---- {f["name"]}')
            if p['reds']:
                open_block()
                for slot, code in p['reds']:
                    w.append(f'            red[{self.dyn_index[slot]}] += '
                             f'act ? ({code}) : 0.0;')
                close_block()
            if pi == red_after:
                w.append('    if (last_item) {')
                w += reduce_call
                w.append('    }')
            if not p['passes']:
                continue
            open_block()
            if p['lam']:
                s = lay['stor'][('lam', fi)]
                for o in range(f['out_core']):
                    w.append(f'            const double lam_{fi}_{o} = '
                             f'{s["name"]}[cfem::Skew<{s["core"]}>::row(row) + {o}];')
            for ps in p['passes']:
                wboff = 0
                staged = []
                for n, it in enumerate(ps):
                    w.append('            {')
                    if it.c == 1:
                        val = it.codes[0]
                        if it.mult:
                            val = f'({it.mult[0]}) * ({val})'
                        w.append(f'                cfem::lane_store(({it.dest})'
                                 f' + kw, lane, nvalid, {val});')
                    else:
                        w.append(f'                double o[{it.c}];')
                        for j, code in enumerate(it.codes):
                            if it.mult:
                                code = f'({it.mult[j]}) * ({code})'
                            w.append(f'                o[{j}] = {code};')
                        w.append(f'                cfem::warp_put<{it.c}>'
                                 f'(wb + {wboff}, lane, o);')
                        staged.append((it, wboff))
                        wboff += _skew_size(it.c, 32)
                    w.append('            }')
                if staged:
                    w.append('            __syncwarp();')
                    for it, wboff in staged:
                        w.append(f'            cfem::warp_flush<{it.c}>(wb + '
                                 f'{wboff}, lane, ({it.dest}) + kw * {it.c}, '
                                 'nvalid);')
                    w.append('            __syncwarp();')
            close_block()
        w.append('    if (!last_item) __syncthreads();   // staging buffer '
                 'is refilled by the next prefetch')
        w.append('    }   // item loop')
        if has_red and red_after < 0:
            # reductions without a per-sample term still go through the tree
            w.append('    {')
            w += reduce_call
            w.append('    }')
        w.append('    // programmatic dependent launch: the parameter-only '
                 'kernel is the prerequisite grid;')
        w.append('    // nothing here reads its results, but a completed '
                 'per-sample kernel must imply a')
        w.append('    // completed (and flushed) parameter-only kernel for '
                 'whatever follows in the stream')
        w.append('    asm volatile("griddepcontrol.wait;" ::: "memory");')
        w.append('}')
        return '\n'.join(w), lay['total'] * 8

    def _param_entries(self):
        """Every entry of the parameter-only functions, grouped G, JAC, HESS:
        dicts with the function, its dependencies, the destination (array
        kind, block, offset), the value code and the multiplier of Hessian
        entries (a constraint multiplier or obj_factor)."""
        st = self.st
        ents = {G: [], JAC: [], HESS: []}
        for fi in self.param_funs:
            f = self.funs[fi]
            spec = f['spec']
            if f['is_objective']:
                continue
            ci = f['cons_index']
            for e in spec.values:
                ents[G].append(dict(fun=f, deps=e.deps, kind=0, blk=ci,
                                    off=e.index[0], code=e.code, mult=None))
        for bi, blk in enumerate(st.jac_blocks):
            f = self.funs[blk['fun']]
            if f['per_sample']:
                continue
            for j, e in enumerate(blk['entries']):
                ents[JAC].append(dict(fun=f, deps=e.deps, kind=1, blk=bi,
                                      off=j, code=e.code, mult=None))
        for bi, blk in enumerate(st.hess_blocks):
            f = self.funs[blk['fun']]
            if f['per_sample']:
                continue
            for j, e in enumerate(blk['entries']):
                mult = (('sigma',) if f['is_objective'] else
                        ('lam', f['cons_index'], e.index[2]))
                ents[HESS].append(dict(fun=f, deps=e.deps, kind=2, blk=bi,
                                       off=j, code=e.code, mult=mult))
        return ents

    _DEST = ('a.g[b * a.ncons + a.cons_off[{blk}] + {off}]',
             'a.jac[b * a.nnz_jac + a.jac_off[{blk}] + {off}]',
             'a.hess[b * a.nnz_hess + a.hess_off[{blk}] + {off}]')

    @staticmethod
    def _mult_code(mult):
        if mult is None:
            return None
        if mult[0] == 'sigma':
            return 'a.obj_factor'
        return f'lam[a.cons_off[{mult[1]}] + {mult[2]}]'

    @staticmethod
    def _split_top(text, seps):
        """Split ``text`` at the top-level (outside parentheses) occurrences
        of the separators; returns (pieces, separators)."""
        pieces, used, depth, cur, i = [], [], 0, [], 0
        while i < len(text):
            ch = text[i]
            if ch == '(':
                depth += 1
            elif ch == ')':
                depth -= 1
            hit = None
            if depth == 0:
                for sp in seps:
                    if text.startswith(sp, i):
                        hit = sp
                        break
            if hit:
                pieces.append(''.join(cur))
                used.append(hit)
                cur = []
                i += len(hit)
            else:
                cur.append(ch)
                i += 1
        pieces.append(''.join(cur))
        return pieces, used

    @classmethod
    def _parse_sum(cls, code, allow_group=True):
        """``code`` as a sum of products in the printed order:
        ``[(coeff, [identifier, ...], inner)]`` where ``inner`` is None or the
        parsed flat sum of ONE trailing parenthesised factor
        (``c*a*b*(s1 + s2 + ...)``, the Horner-like form sympy prints for the
        ZOH discretisation); None if the expression is anything else.
        ``a - b*c`` becomes ``+1*a`` and ``-1*b*c``: sign flips are exact in
        IEEE arithmetic, so evaluating terms and factors left to right
        reproduces the printed expression operation for operation."""
        import re
        num = r'[0-9]+(\.[0-9]*)?([eE][-+]?[0-9]+)?'
        text = code.strip()
        sign = 1.0
        if text.startswith('-'):
            sign, text = -1.0, text[1:].strip()
        bodies, seps = cls._split_top(text, (' + ', ' - '))
        terms = []
        for i, body in enumerate(bodies):
            if i > 0:
                sign = 1.0 if seps[i - 1] == ' + ' else -1.0
            facs, _ = cls._split_top(body.strip(), ('*',))
            coeff, idents, inner = 1.0, [], None
            for k, t in enumerate(facs):
                t = t.strip()
                if inner is not None:
                    return None         # the group must be the last factor
                if re.fullmatch(r'v_[A-Za-z0-9_]+', t):
                    idents.append(t)
                elif k == 0 and re.fullmatch(num, t):
                    coeff = float(t)
                elif k == 0 and re.fullmatch(rf'\(({num})/({num})\)', t):
                    a_, b_ = t[1:-1].split('/')
                    coeff = float(a_) / float(b_)   # folded by the C compiler too
                elif (allow_group and t.startswith('(') and t.endswith(')')
                      and k == len(facs) - 1 and k > 0):
                    inner = cls._parse_sum(t[1:-1], allow_group=False)
                    if inner is None:
                        return None
                else:
                    return None
            terms.append((sign * coeff, idents, inner))
        return terms

    @classmethod
    def _parse_flat(cls, code):
        terms = cls._parse_sum(code, allow_group=False)
        return None if terms is None else [(c, ids) for c, ids, _ in terms]

    #: entries per __noinline__ chunk function of the parameter kernel (ptxas
    #: time grows super-linearly with the size of one function)
    PARAM_CHUNK = 32

    def _pack_factor(self, fun, arg, flat):
        """Operand of a table term: space (0 decision variable, 1 scalar),
        index of the variable / scalar, element."""
        ref = fun['args'][arg]
        space = {'param': 0, 'scalar': 1}[ref[0]]
        assert 0 <= ref[1] < 1024 and 0 <= flat < (1 << 20)
        return (space << 30) | (ref[1] << 20) | (flat if space == 0 else 0)

    def _emit_param_kernel(self):
        """The parameter-only constraints are polynomials of the parameters.
        Entries that print as a flat sum of products (over 90 % of them) go
        into a TABLE -- per entry a destination, a term range and an optional
        multiplier (lambda / obj_factor); per term a coefficient, a factor
        range and optionally the term range of ONE trailing parenthesised
        factor (the Horner-like form of the ZOH discretisation) -- evaluated
        by one uniform loop, a thread per (entry, problem):
        no per-entry code, no divergence, compile time and instruction-cache
        footprint independent of the number of entries.  The rest (nested,
        Horner-like expressions of the ZOH discretisation) keeps generated
        straight-line code, one WARP per entry."""
        ents = self._param_entries()
        order = ents[G] + ents[JAC] + ents[HESS]
        self.n_param_entries = len(order)
        table, code_ents = [], []
        for ent in order:
            terms = self._parse_sum(ent['code'])
            if terms is None:
                code_ents.append(ent)
                continue
            idmap = {symoptim.c_ident(a_, fl): (a_, fl) for a_, fl in ent['deps']}
            used = [i for _, ids, inner in terms
                    for i in ids + [j for _, jds, _ in (inner or [])
                                    for j in jds]]
            if any(i not in idmap for i in used):
                code_ents.append(ent)
                continue

            def pack(ids, ent=ent, idmap=idmap):
                return [self._pack_factor(ent['fun'], *idmap[i]) for i in ids]
            ent['terms'] = [(c, pack(ids), None if inner is None else
                             [(ci, pack(jds)) for ci, jds, _ in inner])
                            for c, ids, inner in terms]
            table.append(ent)
        self.n_param_table, self.n_param_code = len(table), len(code_ents)
        CH = self.PARAM_CHUNK
        w = []
        # ---- tables
        # terms of an entry are contiguous; the inner sums (terms of a
        # trailing parenthesised factor) are stored after all entry terms
        term0, coeffs, fac0, facs, dests, mults = [0], [], [0], [], [], []
        inner_of = []           # per term: None or list of inner terms
        for ent in table:
            for c, fl, inner in ent['terms']:
                coeffs.append(c)
                facs += fl
                fac0.append(len(facs))
                inner_of.append(inner)
            term0.append(len(coeffs))
        in0, in1 = [], []
        for inner in list(inner_of):
            if inner is None:
                in0.append(-1)
                in1.append(-1)
                continue
            in0.append(len(coeffs))
            for c, fl in inner:
                coeffs.append(c)
                facs += fl
                fac0.append(len(facs))
            in1.append(len(coeffs))
        for ent in table:
            assert ent['blk'] < 1024 and ent['off'] < (1 << 20)
            dests.append((ent['kind'] << 30) | (ent['blk'] << 20) | ent['off'])
            m = ent['mult']
            if m is None:
                mults.append(0xFFFFFFFF)
            elif m[0] == 'sigma':
                mults.append(0xFFFFFFFE)
            else:
                assert m[1] < 1024 and m[2] < (1 << 20)
                mults.append((m[1] << 20) | m[2])

        # kept for inspection / the CPU test of the table builder
        self.param_table = dict(
            entries=table, term0=term0, coeff=coeffs, fac0=fac0, fac=facs,
            in0=in0, in1=in1, dest=dests, mult=mults)

        def arr(ctype, name, vals, fmt):
            body = ', '.join(fmt(v) for v in vals) if vals else fmt(0)
            lines = [f'static __device__ const {ctype} {name}'
                     f'[{max(1, len(vals))}] = {{']
            items = body.split(', ')
            for i in range(0, len(items), 12):
                lines.append('    ' + ', '.join(items[i:i + 12]) + ',')
            lines.append('};')
            return lines
        w.append(f'constexpr int kParamTable = {len(table)};   '
                 '// entries evaluated from the tables')
        w.append(f'constexpr int kParamCode = {len(code_ents)};    '
                 '// entries with generated code')
        w += arr('int', 'kPE_term0', term0, str)
        w += arr('unsigned', 'kPE_dest', dests, lambda v: f'{v}u')
        w += arr('unsigned', 'kPE_mult', mults, lambda v: f'{v}u')
        w += arr('double', 'kPT_coeff', coeffs, lambda v: repr(float(v)))
        w += arr('int', 'kPT_fac0', fac0, str)
        w += arr('int', 'kPT_in0', in0, str)
        w += arr('int', 'kPT_in1', in1, str)
        w += arr('unsigned', 'kPF', facs, lambda v: f'{v}u')
        # ---- generated code for the remaining entries
        nchunks = (len(code_ents) + CH - 1) // CH
        for c in range(nchunks):
            w.append(f'static __device__ __noinline__ void cfem_param_chunk{c}('
                     'const cfem::KArgs& a, const long long b, const int e)')
            w.append('{')
            w.append('    const double* __restrict__ dvec = a.dvec + b * a.ndec;')
            w.append('    const double* __restrict__ lam = a.lam + b * a.ncons;')
            w.append('    (void)dvec; (void)lam;')
            w.append('    switch (e) {')
            for i in range(c * CH, min(len(code_ents), (c + 1) * CH)):
                ent = code_ents[i]
                dest = self._DEST[ent['kind']].format(**ent)
                mult = self._mult_code(ent['mult'])
                code = ent['code'] if mult is None else \
                    f'({mult}) * ({ent["code"]})'
                w.append(f'    case {i}: {{')
                defs, undefs = self._define_args(ent['fun'], ent['deps'],
                                                 'global')
                w += defs
                w.append(f'        if (mask & {(G, JAC, HESS)[ent["kind"]]}u) '
                         f'{dest} = {code};')
                w += undefs
                w.append('        break; }')
            w.append('    default: break;')
            w.append('    }')
            w.append('}')
        for c in range(nchunks):    # the chunks take the mask
            pass
        text = '\n'.join(w).replace(
            'const cfem::KArgs& a, const long long b, const int e)',
            'const cfem::KArgs& a, const unsigned mask, const long long b, '
            'const int e)')
        w = [text]
        w.append('// Launched right before the per-sample kernel on the same '
                 'stream; the latter is')
        w.append('// launched with programmatic stream serialisation and never '
                 'waits on this grid')
        w.append('// (there is no data dependence), so the two overlap without '
                 'a second stream or')
        w.append('// fork/join events: this kernel releases its dependents '
                 'first thing.')
        w.append(f'constexpr int kParamTableCtas = ({len(table)} + 127) / 128;')
        w.append('__global__ void __launch_bounds__(128)')
        w.append('cfem_param_kernel(const __grid_constant__ cfem::KArgs a, '
                 'const unsigned mask, const int batch)')
        w.append('{')
        w.append('    asm volatile("griddepcontrol.launch_dependents;");')
        w.append('    const long long b = blockIdx.y;      // problem of a batch')
        w.append('    (void)batch;')
        w.append('    if ((int)blockIdx.x < kParamTableCtas) {')
        w.append('        // table entries: a thread per entry, one uniform loop')
        w.append('        const int e = blockIdx.x * 128 + threadIdx.x;')
        w.append('        if (e >= kParamTable) return;')
        w.append('        const unsigned dest = kPE_dest[e];')
        w.append('        const unsigned kind = dest >> 30;')
        w.append(f'        if (!(mask & ({G}u << kind))) return;')
        w.append('        const double* __restrict__ dvec = a.dvec + b * a.ndec;')
        w.append('        // coefficient times factors, left to right (the printed '
                 'order, no contraction)')
        w.append('        auto product = [&](int t) {')
        w.append('            double p = kPT_coeff[t];')
        w.append('            for (int k = kPT_fac0[t]; k < kPT_fac0[t + 1]; ++k) {')
        w.append('                const unsigned f = kPF[k];')
        w.append('                const double v = (f >> 30) == 0u ? '
                 'dvec[a.var_off[(f >> 20) & 1023u] + (f & 0xFFFFFu)] '
                 ': a.scalars[(f >> 20) & 1023u];')
        w.append('                p = __dmul_rn(p, v);')
        w.append('            }')
        w.append('            return p;')
        w.append('        };')
        w.append('        double sum = 0.0;')
        w.append('        for (int t = kPE_term0[e]; t < kPE_term0[e + 1]; ++t) {')
        w.append('            double p = product(t);')
        w.append('            if (kPT_in0[t] >= 0) {      // ... * (inner sum)')
        w.append('                double inner = 0.0;')
        w.append('                for (int u = kPT_in0[t]; u < kPT_in1[t]; ++u) '
                 'inner = __dadd_rn(inner, product(u));')
        w.append('                p = __dmul_rn(p, inner);')
        w.append('            }')
        w.append('            sum = __dadd_rn(sum, p);')
        w.append('        }')
        w.append('        const unsigned m = kPE_mult[e];')
        w.append('        if (m == 0xFFFFFFFEu) sum = __dmul_rn(a.obj_factor, sum);')
        w.append('        else if (m != 0xFFFFFFFFu) sum = __dmul_rn('
                 'a.lam[b * a.ncons + a.cons_off[m >> 20] + (m & 0xFFFFFu)], sum);')
        w.append('        const long long blk = (dest >> 20) & 1023u, '
                 'off = dest & 0xFFFFFu;')
        w.append('        if (kind == 0u) a.g[b * a.ncons + a.cons_off[blk] + off] = sum;')
        w.append('        else if (kind == 1u) a.jac[b * a.nnz_jac + a.jac_off[blk] + off] = sum;')
        w.append('        else a.hess[b * a.nnz_hess + a.hess_off[blk] + off] = sum;')
        w.append('        return;')
        w.append('    }')
        w.append('    // generated entries: each has its own straight-line code, '
                 'ONE WARP per entry')
        w.append('    // (a thread per entry would make every warp 32-way divergent)')
        w.append('    const int e = ((int)blockIdx.x - kParamTableCtas) * 4 + '
                 '(threadIdx.x >> 5);')
        w.append('    if (e >= kParamCode || (threadIdx.x & 31)) return;')
        w.append(f'    switch (e / {CH}) {{')
        for c in range(nchunks):
            w.append(f'    case {c}: cfem_param_chunk{c}(a, mask, b, e); break;')
        w.append('    default: break;')
        w.append('    }')
        w.append('}')
        return '\n'.join(w)

    def _emit_finalize(self):
        R = len(self.slots)
        nd = max(1, len(self.dyn_slots))
        w = []
        w.append('// Objective / parameter-gradient entries from the reduction slots.')
        w.append('static __device__ __forceinline__ void cfem_write_sums('
                 f'const cfem::KArgs& a, const unsigned mask, const long long b, '
                 f'const double (&tot)[{R}])')
        w.append('{')
        w.append(f'    if (mask & {F}u) a.f[b] = tot[0];')
        w.append(f'    if (mask & {GRAD}u) {{')
        for i, sl in enumerate(self.slots[1:], start=1):
            w.append(f'        a.grad[b * a.ndec + a.var_off[{sl["var"]}] + '
                     f'{sl["flat"]}] = tot[{i}];')
        w.append('    }')
        w.append('}')
        w.append('')
        w.append('// Runs in the last CTA of a problem (all CFEM_TILE threads).')
        w.append('static __device__ __noinline__ void cfem_finalize('
                 'const cfem::KArgs& a, const unsigned mask, const long long b, '
                 'const int tid, double* scratch)')
        w.append('{')
        w.append('    const double* __restrict__ dvec = a.dvec + b * a.ndec;')
        w.append(f'    double tot[{R}];')
        w.append(f'    for (int r = 0; r < {R}; ++r) tot[r] = 0.0;')
        if self.retire == 'flag':
            # a structure without per-sample reduction terms still publishes
            # one (zero) slot per CTA: the finaliser must see every CTA
            for di, slot in enumerate(self.dyn_slots or [None]):
                call = (f'cfem::collect_partials(a, b, {nd}, {di}, scratch, '
                        'tid);')
                w.append(f'    tot[{slot}] = {call}' if slot is not None
                         else f'    (void){call}')
        else:
            w.append(f'    const double* part = a.gpartials + b * '
                     f'a.group_stride * {nd};')
            for di, slot in enumerate(self.dyn_slots):
                w.append(f'    tot[{slot}] = cfem::reduce_tiles<CFEM_TILE>(part, '
                         f'a.ngroups, {nd}, {di}, scratch, tid);')
        w.append('    if (tid >= 32) return;      // warp 0 goes on; '
                 'the sums are in thread 0')
        w.append('    if (tid == 0) {')
        for fi, f in enumerate(self.funs):
            if not f['is_objective']:
                continue
            spec = f['spec']
            rows = f'(double)a.fun_rows[{fi}]'
            const = []      # (slot, code, deps)
            for e in spec.values:
                for code, d in e.terms:
                    if not (f['per_sample'] and self._is_sample_dep(f, d)):
                        const.append((0, code, d))
            for wrt in spec.wrt:
                ref = f['args'][wrt]
                if ref[0] != 'param':
                    continue
                for e in spec.jac.get(wrt, []):
                    slot = self.slot_of[(ref[1], e.index[0])]
                    for code, d in e.terms:
                        if not (f['per_sample']
                                and self._is_sample_dep(f, d)):
                            const.append((slot, code, d))
            if not const:
                continue
            deps = sorted({x for _, _, d in const for x in d})
            defs, undefs = self._define_args(f, deps, 'global')
            w.append(f'    // sample-independent terms of {f["name"]}')
            w += defs
            for slot, code, _ in const:
                w.append(f'    tot[{slot}] += {rows} * ({code});')
            w += undefs
        w.append('    // this rank\'s own sums, double-buffered by launch parity '
                 '(the side-stream exchange kernel')
        w.append('    // of this launch reads them while the next launch is '
                 'already running)')
        w.append(f'    for (int r = 0; r < {R}; ++r) '
                 f'a.reduce[((a.peer_epoch & 1ull) * gridDim.y + b) * {R} + r] = tot[r];')
        w.append('    }')
        w.append('    // time-sharded run: exchange the partial sums with the peer '
                 'GPUs through NVLink-mapped')
        w.append('    // memory, inside this kernel; lane p of this warp talks '
                 'to rank p')
        w.append('    if (a.peer_world > 1) {')
        w.append(f'        for (int r = 0; r < {R}; ++r) '
                 'tot[r] = __shfl_sync(0xffffffffu, tot[r], 0);')
        w.append('        if (a.peer_defer) {')
        w.append('            // pipelined mode: finish the PREVIOUS launch, post this one')
        w.append('            if (a.peer_epoch > 1ull && a.peer_prev_mask) {')
        w.append(f'                double prev[{R}];')
        w.append(f'                cfem::peer_collect<{R}>(a, b, (long long)gridDim.y, '
                 'a.peer_epoch - 1ull, prev, tid);')
        w.append('                if (tid == 0) cfem_write_sums(a, a.peer_prev_mask, '
                 'b, prev);')
        w.append('            }')
        w.append(f'            cfem::peer_post<{R}>(a, b, (long long)gridDim.y, '
                 'tot, tid);')
        w.append('            return;')
        w.append('        }')
        w.append(f'        cfem::peer_allreduce<{R}>(a, b, tot, tid);')
        w.append('    }')
        w.append('    if (tid == 0) cfem_write_sums(a, mask, b, tot);')
        w.append('}')
        w.append('')
        w.append('__global__ void cfem_apply_reduced_kernel(const cfem::KArgs a, '
                 'const double* __restrict__ red)')
        w.append('{')
        w.append('    const long long b = blockIdx.x;')
        w.append('    if (threadIdx.x != 0) return;')
        w.append(f'    a.f[b] = red[b * {R}];')
        for i, s in enumerate(self.slots[1:], start=1):
            w.append(f'    a.grad[b * a.ndec + a.var_off[{s["var"]}] + '
                     f'{s["flat"]}] = red[b * {R} + {i}];')
        w.append('}')
        w.append('')
        w.append('// Pipelined cross-GPU reduction, exchange half, on a side '
                 'stream right after the')
        w.append('// per-sample kernel of launch a.peer_epoch (which ran WITHOUT '
                 'peers and left this')
        w.append('// rank\'s sums in a.reduce): finish the previous launch, then '
                 'post this one.  It')
        w.append('// overlaps the next per-sample kernel, so the NVLink round '
                 'trips and the system')
        w.append('// fence are in nobody\'s way.')
        w.append('__global__ void cfem_peer_exchange_kernel(const cfem::KArgs a)')
        w.append('{')
        w.append('    const long long b = blockIdx.x, nb = gridDim.x;')
        w.append('    const int lane = (int)threadIdx.x;')
        w.append(f'    double tot[{R}];')
        w.append(f'    for (int r = 0; r < {R}; ++r) '
                 f'tot[r] = __ldcg(a.reduce + ((a.peer_epoch & 1ull) * nb + b) * {R} + r);')
        w.append('    if (a.peer_epoch > 1ull && a.peer_prev_mask) {')
        w.append(f'        double prev[{R}];')
        w.append(f'        cfem::peer_collect<{R}>(a, b, nb, a.peer_epoch - 1ull, '
                 'prev, lane);')
        w.append('        if (lane == 0) cfem_write_sums(a, a.peer_prev_mask, b, prev);')
        w.append('    }')
        w.append(f'    cfem::peer_post<{R}>(a, b, nb, tot, lane);')
        w.append('}')
        w.append('')
        w.append('// Pipelined cross-GPU reduction: waits for the posts of '
                 'a.peer_epoch from all')
        w.append('// ranks, sums them in rank order and writes the objective / '
                 'parameter gradient.')
        w.append('__global__ void cfem_peer_collect_kernel(const cfem::KArgs a, '
                 'const unsigned mask)')
        w.append('{')
        w.append('    const long long b = blockIdx.x;')
        w.append(f'    double tot[{R}];')
        w.append(f'    cfem::peer_collect<{R}>(a, b, (long long)gridDim.x, '
                 'a.peer_epoch, tot, (int)threadIdx.x);')
        w.append('    if (threadIdx.x == 0) cfem_write_sums(a, mask, b, tot);')
        w.append('}')
        return '\n'.join(w)

    def model_json(self):
        st = self.st
        return {
            'abi': 1, 'tile': self.tile, 'masks': list(self.masks),
            'vars': st.vars,
            'data': [{k: d[k] for k in ('name', 'core', 'r0', 'hshift')}
                     for d in st.data],
            'scalars': st.scalars,
            'funs': [{k: f[k] for k in ('name', 'is_objective', 'per_sample',
                                        'r0', 'out_core', 'cons_index')}
                     for f in st.funs],
            'jac_blocks': [{'fun': b['fun'], 'wrt': b['wrt'], 'c': b['c']}
                           for b in st.jac_blocks],
            'hess_blocks': [{'fun': b['fun'], 'pair': list(b['pair']),
                             'c': b['c']} for b in st.hess_blocks],
            'reduce': [[s['var'], s['flat']] for s in self.slots],
        }

    def _tables(self):
        """Structure tables of namespace ``gen`` (shared by both units)."""
        st = self.st
        ncons = sum(1 for f in st.funs if not f['is_objective'])

        def table(ctype, name, rows):
            body = ',\n'.join('    ' + r for r in rows) if rows else '    {}'
            return (f'constexpr {ctype} {name}[{max(1, len(rows))}] = '
                    f'{{\n{body}\n}};')

        w = []
        w.append('namespace gen {')
        w.append('struct VarDesc { const char* name; int core; int per_sample; '
                 'int r0; int hshift; };')
        w.append('struct DataDesc { const char* name; int core; int r0; '
                 'int hshift; };')
        w.append('struct FunDesc { const char* name; int is_objective; '
                 'int per_sample; int r0; int out_core; int cons_index; };')
        w.append('struct BlockDesc { int fun; int c; };')
        w.append(f'constexpr int kNumVars = {len(st.vars)};')
        w.append(f'constexpr int kNumData = {len(st.data)};')
        w.append(f'constexpr int kNumScalars = {len(st.scalars)};')
        w.append(f'constexpr int kNumFuns = {len(st.funs)};')
        w.append(f'constexpr int kNumCons = {ncons};')
        w.append(f'constexpr int kNumJacBlocks = {len(st.jac_blocks)};')
        w.append(f'constexpr int kNumHessBlocks = {len(st.hess_blocks)};')
        w.append(f'constexpr int kNumReduce = {len(self.slots)};')
        w.append(f'constexpr int kNumDynReduce = '
                 f'{max(1, len(self.dyn_slots))};')
        w.append(f'constexpr int kNumMasks = {len(self.masks)};')
        w.append(table('VarDesc', 'kVars', [
            f'{{"{v["name"]}", {v["core"]}, {v["per_sample"]}, {v["r0"]}, '
            f'{v["hshift"]}}}' for v in st.vars]))
        w.append(table('DataDesc', 'kData', [
            f'{{"{d["name"]}", {d["core"]}, {d["r0"]}, {d["hshift"]}}}'
            for d in st.data]))
        w.append(table('FunDesc', 'kFuns', [
            f'{{"{f["name"]}", {f["is_objective"]}, {f["per_sample"]}, '
            f'{f["r0"]}, {f["out_core"]}, {f["cons_index"]}}}'
            for f in st.funs]))
        w.append(table('BlockDesc', 'kJacBlocks', [
            f'{{{b["fun"]}, {b["c"]}}}' for b in st.jac_blocks]))
        w.append(table('BlockDesc', 'kHessBlocks', [
            f'{{{b["fun"]}, {b["c"]}}}' for b in st.hess_blocks]))
        w.append('constexpr unsigned kMasks[] = {'
                 + ', '.join(f'{m}u' for m in self.masks) + '};')
        w.append('}  // namespace gen')
        return w

    def sources(self):
        """``{'main': ..., 'param': ...}``: the two translation units of one
        model library.  ``param`` holds only the parameter-only constraint
        kernel (long straight-line code, slow to compile, depends on nothing
        hand-written but cfem_args.cuh); ``main`` holds the per-sample
        kernels, the reductions and the host side of the C ABI."""
        st = self.st
        kernels, smem = [], {}
        param_kernel = self._emit_param_kernel()    # sets n_param_entries
        for m in self.masks:
            text, nbytes = self._emit_sample_kernel(m)
            kernels.append(text)
            smem[m] = nbytes
        finalize = self._emit_finalize()
        mj = json.dumps(self.model_json(), sort_keys=True)
        cstr = '\n'.join('    "' + mj[i:i + 100].replace('\\', '\\\\')
                         .replace('"', '\\"') + '"'
                         for i in range(0, len(mj), 100))
        head = ['// GENERATED by colloc_fem_code_b200/codegen.py -- do not edit.',
                '// Model structure: ' + ', '.join(f['name'] for f in st.funs),
                '#include <cuda_runtime.h>', '#include <math.h>']
        tables = self._tables()
        launch_param_sig = ('cudaError_t launch_param(unsigned mask, int batch, '
                            'cudaStream_t s, const cfem::KArgs& a)')

        # ---- parameter-only unit
        w = list(head) + tables
        w.append('#include "cfem_args.cuh"')
        w.append('namespace gen {')
        w.append(param_kernel)
        w.append(launch_param_sig)
        w.append('{')
        if self.n_param_entries:
            w.append('    const dim3 grid(kParamTableCtas + (kParamCode + 3) / 4, '
                     'batch);')
            w.append('    cfem_param_kernel<<<grid, 128, 0, s>>>(a, mask, batch);')
            w.append('    return cudaGetLastError();')
        else:
            w.append('    (void)mask; (void)batch; (void)s; (void)a;')
            w.append('    return cudaSuccess;')
        w.append('}')
        w.append('}  // namespace gen')
        param_unit = '\n'.join(w) + '\n'

        # ---- main unit
        w = list(head)
        if self.experiment:
            w.append(f'#define CFEM_EXPERIMENT_{self.experiment.upper()} 1')
        if self.store:
            op = {'wb': '(*(ptr) = (val))', 'cg': '__stcg((ptr), (val))',
                  'wt': '__stwt((ptr), (val))'}[self.store]
            w.append(f'#define CFEM_STORE_OP(ptr, val) {op}')
        w.append(f'#define CFEM_TILE {self.tile}')
        w += tables
        w.append('namespace gen {')
        w.append('const char kModelJson[] =\n' + cstr + ';')
        w.append('}  // namespace gen')
        w.append('#include "cfem_device.cuh"')
        w.append('namespace gen {')
        w.append(finalize)
        w += kernels
        w.append(f'constexpr int kNumParamEntries = {self.n_param_entries};')
        w.append('// partial sums retire through self-validating slots and a '
                 'finaliser CTA')
        w.append('constexpr bool kFlagRetire = '
                 f'{"true" if self.retire == "flag" else "false"};')
        w.append(launch_param_sig + ';    // parameter-only unit')
        w.append(f'static int g_ctas_per_sm[{len(self.masks)}];')
        w.append(f'static int g_ctas_per_sm1[{len(self.masks)}];   '
                 '// with the single-buffer shared-memory size')
        w.append('static bool g_single_buf = true;      // CFEM_SINGLE_BUF=0: '
                 'always both staging buffers')
        w.append('static bool g_finaliser = true;       // CFEM_FINALISER=0: the '
                 'last tile CTA finalises (time-sharded runs)')
        w.append('static const size_t kSmem2[] = {'
                 + ', '.join(f'kSmemBytes_m{m}' for m in self.masks) + '};')
        w.append('static const size_t kSmem1[] = {'
                 + ', '.join(f'kSmemBytes1_m{m}' for m in self.masks) + '};')
        w.append('static cudaError_t configure_kernels()')
        w.append('{')
        w.append('    cudaError_t e = cudaSuccess;')
        for i, m in enumerate(self.masks):
            if smem[m] > 48 * 1024:
                w.append(f'    e = cudaFuncSetAttribute(cfem_sample_kernel_m{m}, '
                         'cudaFuncAttributeMaxDynamicSharedMemorySize, '
                         f'(int)kSmemBytes_m{m});')
                w.append('    if (e != cudaSuccess) return e;')
            w.append('    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor('
                     f'&g_ctas_per_sm[{i}], cfem_sample_kernel_m{m}, CFEM_TILE, '
                     f'kSmemBytes_m{m});')
            w.append('    if (e != cudaSuccess) return e;')
            w.append(f'    if (g_ctas_per_sm[{i}] < 1) g_ctas_per_sm[{i}] = 1;')
            w.append('    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor('
                     f'&g_ctas_per_sm1[{i}], cfem_sample_kernel_m{m}, CFEM_TILE, '
                     f'kSmemBytes1_m{m});')
            w.append('    if (e != cudaSuccess) return e;')
            w.append(f'    if (g_ctas_per_sm1[{i}] < 1) g_ctas_per_sm1[{i}] = 1;')
        w.append('    if (const char* v = getenv("CFEM_SINGLE_BUF")) '
                 'g_single_buf = atoi(v) != 0;')
        w.append('    if (const char* v = getenv("CFEM_FINALISER")) '
                 'g_finaliser = atoi(v) != 0;')
        w.append('    return e;')
        w.append('}')
        w.append("""// Work items of one launch (cfem_args.cuh).  Full tiles first; when there is
// enough work, the last two resident sets are half and quarter tiles (graded
// tail): the CTAs that run last -- their slots are not refilled -- are short,
// so the grid drains in a fraction of a tile time.  `resident` = CTAs of this
// problem that fit on the GPU at once.
static void build_items(long long resident, int tail_levels, cfem::KArgs& a)
{
    const long long T = CFEM_TILE, N = a.N;
    const int min_size = 8 * (CFEM_TILE / 32);
    int sizes[cfem::kMaxPhases];
    int np = 1;
    sizes[0] = CFEM_TILE;
    while (np < cfem::kMaxPhases && np <= tail_levels && sizes[np - 1] / 2 >= min_size) {
        sizes[np] = sizes[np - 1] / 2;
        ++np;
    }
    long long tail = 0;
    for (int p = 1; p < np; ++p) tail += resident * sizes[p];
    if (np == 1 || N < tail + 2 * resident * T) {      // not worth grading
        a.nphase = 1;
        a.ph_item0[0] = 0; a.ph_k0[0] = 0; a.ph_size[0] = CFEM_TILE;
        a.nitems = (N + T - 1) / T;
        return;
    }
    a.nphase = np;
    long long item = 0, k0 = 0;
    for (int p = 0; p < np; ++p) {
        a.ph_item0[p] = item; a.ph_k0[p] = k0; a.ph_size[p] = sizes[p];
        long long count = p == 0 ? (N - tail) / T : resident;
        if (p == np - 1) count = (N - k0 + sizes[p] - 1) / sizes[p];   // the rest
        item += count;
        k0 += count * sizes[p];
    }
    a.nitems = item;
}
// Launch geometry: at most `waves` resident sets of CTAs per problem, CTA c
// takes items c, c + nctas, ...  The CTA count is the same for every kernel
// variant of the library (the smallest residency), so the fixed-order
// reductions give the same bits whichever variant serves a callback.  When
// every CTA has exactly one item the second staging buffer is not allocated:
// less shared memory per CTA, more resident CTAs per SM.
static void prepare_sample(unsigned mask, int batch, int sm_count, int waves, int tail_levels,
                           cfem::KArgs& a, dim3& grid, size_t& smem)
{
    int idx = 0, per_sm1 = 1 << 30, per_sm2 = 1 << 30;
    for (int i = 0; i < kNumMasks; ++i) {
        if (kMasks[i] == mask) idx = i;
        if (g_ctas_per_sm1[i] < per_sm1) per_sm1 = g_ctas_per_sm1[i];
        if (g_ctas_per_sm[i] < per_sm2) per_sm2 = g_ctas_per_sm[i];
    }
    long long resident = (long long)per_sm1 * sm_count / batch;
    if (resident < 1) resident = 1;
    build_items(resident, tail_levels, a);
    long long gx = resident * waves;
    smem = kSmem1[idx];
    if (!g_single_buf || gx < a.nitems) {       // several items per CTA: double buffering
        resident = (long long)per_sm2 * sm_count / batch;
        if (resident < 1) resident = 1;
        build_items(resident, tail_levels, a);
        gx = resident * waves;
        smem = kSmem2[idx];
    }
    if (gx > a.nitems) gx = a.nitems;
    if (gx > a.part_stride) gx = a.part_stride;     // partial-sum slots (cfem_create)
    a.nctas = gx;
    a.ngroups = (gx + cfem::kReduceGroup - 1) / cfem::kReduceGroup;
    a.finaliser = ((mask & 3u) && (kFlagRetire || (a.peer_world > 1 && g_finaliser))) ? 1 : 0;
    grid = dim3((unsigned)(gx + a.finaliser), (unsigned)batch);
}""")
        w.append('// overlap_prev: programmatic stream serialisation -- the kernel '
                 'may start while the')
        w.append('// preceding kernel of the stream (the parameter-only kernel, '
                 'no data dependence)')
        w.append('// is still running.')
        w.append('static cudaError_t launch_sample(unsigned mask, int batch, '
                 'int sm_count, int waves, int tail_levels, bool overlap_prev, '
                 'cudaStream_t s, cfem::KArgs a)')
        w.append('{')
        w.append('    cudaLaunchConfig_t cfg = {};')
        w.append('    size_t smem = 0;')
        w.append('    prepare_sample(mask, batch, sm_count, waves, tail_levels, a, cfg.gridDim, smem);')
        w.append('    cfg.dynamicSmemBytes = smem;')
        w.append('    cfg.blockDim = dim3(CFEM_TILE);')
        w.append('    cfg.stream = s;')
        w.append('    cudaLaunchAttribute attr[1];')
        w.append('    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;')
        w.append('    attr[0].val.programmaticStreamSerializationAllowed = 1;')
        w.append('    cfg.attrs = attr;')
        w.append('    cfg.numAttrs = overlap_prev ? 1 : 0;')
        w.append('    switch (mask) {')
        for m in self.masks:
            w.append(f'    case {m}u: '
                     f'return cudaLaunchKernelEx(&cfg, cfem_sample_kernel_m{m}, a);')
        w.append('    default: return cudaErrorInvalidValue;')
        w.append('    }')
        w.append('}')
        w.append('// for CUDA-graph kernel-node identification / updates')
        w.append('static const void* sample_kernel_func(unsigned mask)')
        w.append('{')
        w.append('    switch (mask) {')
        for m in self.masks:
            w.append(f'    case {m}u: return (const void*)cfem_sample_kernel_m{m};')
        w.append('    default: return nullptr;')
        w.append('    }')
        w.append('}')
        w.append('static cudaError_t launch_apply_reduced(int batch, '
                 'cudaStream_t s, const cfem::KArgs& a, const double* red)')
        w.append('{')
        w.append('    cfem_apply_reduced_kernel<<<batch, 32, 0, s>>>(a, red);')
        w.append('    return cudaGetLastError();')
        w.append('}')
        w.append('static cudaError_t launch_peer_exchange(int batch, '
                 'cudaStream_t s, const cfem::KArgs& a)')
        w.append('{')
        w.append('    cfem_peer_exchange_kernel<<<batch, 32, 0, s>>>(a);')
        w.append('    return cudaGetLastError();')
        w.append('}')
        w.append('static cudaError_t launch_peer_collect(unsigned mask, int batch, '
                 'cudaStream_t s, const cfem::KArgs& a)')
        w.append('{')
        w.append('    cfem_peer_collect_kernel<<<batch, 32, 0, s>>>(a, mask);')
        w.append('    return cudaGetLastError();')
        w.append('}')
        w.append('}  // namespace gen')
        w.append('#include "cfem_host.inl"')
        return {'main': '\n'.join(w) + '\n', 'param': param_unit}


def generate(structure, **kwargs):
    """CUDA sources ``{'main', 'param'}`` for an ``optim.Structure``."""
    return Generator(structure, **kwargs).sources()
