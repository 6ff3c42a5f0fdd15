"""Problem glue: the ``ceacoest.optim`` boundary of the reference.

/root/reference/fem.py:9-57 builds its problems on ``optim.Problem``:
``add_decision`` (flat decision-vector layout), ``optim.Decision`` views
(``xprev`` / ``xnext`` / ``enprev``, fem.py:47-52), ``add_objective`` /
``add_constraint`` with the broadcast output shape, ``variables(dvec)`` and
``unpack_constraints``; the scripts then call ``problem.ipopt(...)``
(/root/reference/attas_sp_ml.py:153-159).  This module keeps that surface.

Evaluation is NOT done here: ``obj`` / ``obj_grad`` / ``constr`` /
``constr_jac_val`` / ``lag_hess_val`` hand the decision vector to the fused
FP64 CUDA kernels generated for this problem structure (``codegen.py``,
``backend.py``).  There is no NumPy evaluation path; a missing CUDA library is
an error.  Only the sparsity *index* arrays are produced on the host.

COO ordering contract: see DESIGN.md ("Sparsity order") -- functions in
registration order (Hessian: objectives first), decision arguments in signature
order (/root/reference/adfem.py:198-209), sample index slowest inside a
(function, wrt) block (adfem.py:303-328), parameter entries of per-sample
functions repeated once per sample (adfem.py:316-318), Hessian entries
oriented ``row >= col``.
"""

import collections
import contextlib

import numpy as np


class Decision:
    """Shape + offset of a block of the decision (or constraint) vector."""

    def __init__(self, shape, offset):
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        self.offset = int(offset)
        self.size = int(np.prod(self.shape, dtype=np.int64))

    def unpack_from(self, vec):
        """Writable view of this block inside ``vec``."""
        return vec[..., self.offset:self.offset + self.size].reshape(
            vec.shape[:-1] + self.shape)

    def __repr__(self):
        return f'Decision(shape={self.shape}, offset={self.offset})'


Component = Decision


class _Registered:
    """A model function registered on a problem with its broadcast shape."""

    def __init__(self, fun, shape, offset, is_objective):
        self.fun = fun
        self.name = fun.__name__
        self.spec = fun.spec                     # symoptim.FunctionSpec
        self.block = Decision(shape, offset)
        core = tuple(self.spec.out_core)
        shape = self.block.shape
        ext = shape[:len(shape) - len(core)]
        if shape[len(ext):] != core:
            raise ValueError(f'{self.name}: registered shape {shape} does not '
                             f'end with the core output shape {core}')
        if len(ext) > 1:
            raise ValueError(f'{self.name}: only one sample axis is supported')
        self.per_sample = len(ext) == 1
        self.rows = int(ext[0]) if ext else 1
        self.is_objective = is_objective


class Structure:
    """N-independent description of a problem; the code generator's input.

    Built from a concrete problem instance, but every row count is stored
    relative to the sample count N, so one compiled library serves any N.
    """

    def __init__(self, problem):
        regs = list(problem.objectives.values()) \
            + list(problem.constraints.values())
        sample_rows = [r.rows for r in regs if r.per_sample]
        if not sample_rows:
            raise ValueError('problem has no per-sample function')
        self.N = N = max(sample_rows)
        aux = problem.variables(np.zeros(problem.ndec))

        self.var_names = list(problem.decision)
        self.vars = []          # dicts: name core per_sample r0 hshift
        self.data = []          # dicts: name core r0 hshift array
        self.scalars = []       # names
        self.scalar_values = []
        self.funs = []          # dicts, objectives first
        var_index = {n: i for i, n in enumerate(self.var_names)}
        sample_core = {}        # decision name -> core size when per-sample
        pending = collections.OrderedDict()   # data arg name -> (array, core)
        arg_refs = []

        def storage_of_dependent(name, core):
            spec = problem.dependent[name]
            for dname, dspec in problem.decision.items():
                delta = spec.offset - dspec.offset
                if (0 <= delta and spec.offset + spec.size
                        <= dspec.offset + dspec.size and delta % core == 0
                        and dspec.size % core == 0):
                    return dname, delta // core
            raise ValueError(f'dependent variable {name} is not a row-shifted '
                             'view of a decision variable')

        for reg in regs:
            spec = reg.spec
            refs = {}
            for a in spec.args:
                core = spec.core_size(a)
                ncore = len(spec.core[a])
                if a in problem.decision or a in problem.dependent:
                    dspec = problem.decision.get(a) or problem.dependent[a]
                    ext = dspec.shape[:len(dspec.shape) - ncore]
                    if dspec.shape[len(ext):] != tuple(spec.core[a]):
                        raise ValueError(f'{reg.name}: variable {a} has shape '
                                         f'{dspec.shape}, core {spec.core[a]}')
                    if not ext:
                        if a in problem.dependent:
                            raise ValueError('parameter-like dependent '
                                             f'variable {a} is not supported')
                        refs[a] = ('param', var_index[a])
                        continue
                    if not reg.per_sample or ext != (reg.rows,):
                        raise ValueError(f'{reg.name}: argument {a} has rows '
                                         f'{ext}, function has {reg.rows}')
                    if a in problem.decision:
                        dname, shift = a, 0
                    else:
                        dname, shift = storage_of_dependent(a, core)
                    if sample_core.setdefault(dname, core) != core:
                        raise ValueError(f'variable {dname} is viewed with '
                                         'two different core sizes')
                    refs[a] = ('var', var_index[dname], int(shift))
                else:
                    if a not in aux:
                        raise KeyError(f'{reg.name}: no value for argument '
                                       f'{a!r} in problem.variables()')
                    val = aux[a]
                    if np.ndim(val) == ncore:
                        if ncore != 0:
                            raise ValueError(f'auxiliary parameter {a} must '
                                             'be a scalar')
                        if a not in self.scalars:
                            self.scalars.append(a)
                            self.scalar_values.append(float(val))
                        refs[a] = ('scalar', self.scalars.index(a))
                    else:
                        val = np.asarray(val)
                        if not reg.per_sample or val.shape[0] != reg.rows:
                            raise ValueError(f'{reg.name}: data {a} has '
                                             f'{val.shape[0]} rows, function '
                                             f'has {reg.rows}')
                        pending.setdefault(a, (val, core))
                        refs[a] = ('data', a)
            arg_refs.append(refs)

        # data arrays: the largest array of every aliasing group is the root,
        # the others (e.g. uprev = u[:-1], fem.py:23) are row-shifted views
        resolved = {}
        order = sorted(pending, key=lambda n: -pending[n][0].nbytes)
        for name in order:
            arr, core = pending[name]
            ptr = arr.__array_interface__['data'][0]
            for idx, d in enumerate(self.data):
                root = d['source']
                rptr = root.__array_interface__['data'][0]
                if (d['core'] == core and arr.strides == root.strides
                        and np.shares_memory(arr, root) and ptr >= rptr
                        and (ptr - rptr) % root.strides[0] == 0):
                    resolved[name] = (idx, (ptr - rptr) // root.strides[0])
                    break
            else:
                resolved[name] = (len(self.data), 0)
                self.data.append({'name': name, 'core': core, 'source': arr,
                                  'rows': arr.shape[0]})
        for refs in arg_refs:
            for a, ref in refs.items():
                if ref[0] == 'data':
                    refs[a] = ('data', resolved[a][0], int(resolved[a][1]))

        # row counts relative to N, halo need of every storage
        var_shift = collections.defaultdict(int)
        data_shift = collections.defaultdict(int)
        for refs in arg_refs:
            for ref in refs.values():
                if ref[0] == 'var':
                    var_shift[ref[1]] = max(var_shift[ref[1]], ref[2])
                elif ref[0] == 'data':
                    data_shift[ref[1]] = max(data_shift[ref[1]], ref[2])
        for i, name in enumerate(self.var_names):
            dspec = problem.decision[name]
            if name in sample_core:
                core = sample_core[name]
                rows = dspec.size // core
                self.vars.append({'name': name, 'core': core, 'per_sample': 1,
                                  'r0': rows - N, 'hshift': var_shift[i]})
            else:
                self.vars.append({'name': name, 'core': dspec.size,
                                  'per_sample': 0, 'r0': 0, 'hshift': 0})
        for i, d in enumerate(self.data):
            d['r0'] = d['rows'] - N
            d['hshift'] = data_shift[i]

        cons_index = {n: i for i, n in enumerate(problem.constraints)}
        for reg, refs in zip(regs, arg_refs):
            self.funs.append({
                'name': reg.name, 'spec': reg.spec,
                'is_objective': int(reg.is_objective),
                'per_sample': int(reg.per_sample),
                'r0': reg.rows - N if reg.per_sample else 0,
                'out_core': reg.spec.out_size,
                'cons_index': -1 if reg.is_objective
                else cons_index[reg.name],
                'args': refs,
            })

        # derivative blocks in contract order
        self.jac_blocks = []
        self.hess_blocks = []
        for fi, f in enumerate(self.funs):
            spec = f['spec']
            if not f['is_objective']:
                for wrt in spec.wrt:
                    if wrt in spec.jac:
                        self.jac_blocks.append(
                            {'fun': fi, 'wrt': wrt, 'entries': spec.jac[wrt],
                             'c': len(spec.jac[wrt])})
            for pair, entries in spec.hess.items():
                self.hess_blocks.append({'fun': fi, 'pair': pair,
                                         'entries': entries,
                                         'c': len(entries)})

    # -- sizes for a given N (mirrors compute_layout() in cfem_host.inl) -------
    def fun_rows(self, N, halo=0):
        return [N + min(0, f['r0'] + halo) if f['per_sample'] else 1
                for f in self.funs]

    def var_rows(self, N, halo=0):
        return [N + v['r0'] + halo * v['hshift'] if v['per_sample'] else 1
                for v in self.vars]

    def key(self):
        """Hashable, N-independent identity of the structure (cache key)."""
        import json
        funs = []
        for f in self.funs:
            d = {k: v for k, v in f.items() if k != 'spec'}
            d['args'] = {a: list(r) for a, r in f['args'].items()}
            d['spec'] = f['spec'].to_json()
            funs.append(d)
        data = [{k: v for k, v in d.items() if k not in ('source', 'rows')}
                for d in self.data]
        return json.dumps({'vars': self.vars, 'data': data,
                           'scalars': self.scalars, 'funs': funs},
                          sort_keys=True)


class Problem:
    """Drop-in for ``ceacoest.optim.Problem`` backed by CUDA evaluation."""

    def __init__(self):
        self.decision = collections.OrderedDict()
        """Decision variable specifications (fem.py:141 tests membership)."""
        self.dependent = collections.OrderedDict()
        self.objectives = collections.OrderedDict()
        self.constraints = collections.OrderedDict()
        self.ndec = 0
        """Size of the decision vector."""
        self.ncons = 0
        """Size of the constraint vector."""
        self._structure = None
        self._backend = None
        self._index_cache = {}

    # -- registration (fem.py:36-57) -------------------------------------------
    def _invalidate(self):
        if self._backend is not None:
            raise RuntimeError('the problem was already compiled for the GPU; '
                               'register everything before evaluating')
        self._structure = None
        self._index_cache = {}

    def add_decision(self, name, shape):
        self._invalidate()
        spec = Decision(shape, self.ndec)
        self.decision[name] = spec
        self.ndec += spec.size
        return spec

    def add_dependent_variable(self, name, spec):
        self._invalidate()
        self.dependent[name] = spec

    def add_objective(self, fun, shape):
        self._invalidate()
        self.objectives[fun.__name__] = _Registered(fun, shape, 0, True)

    def add_constraint(self, fun, shape):
        self._invalidate()
        reg = _Registered(fun, shape, self.ncons, False)
        self.constraints[fun.__name__] = reg
        self.ncons += reg.block.size

    # -- packing -----------------------------------------------------------------
    def variables(self, dvec):
        """Dict of writable views into ``dvec`` (attas_sp_ml.py:96-111)."""
        dvec = np.asarray(dvec)
        out = {n: s.unpack_from(dvec) for n, s in self.decision.items()}
        out.update((n, s.unpack_from(dvec)) for n, s in self.dependent.items())
        return out

    def unpack_constraints(self, cvec):
        """Dict of writable views by constraint name (attas_sp_ml.py:137-141)."""
        cvec = np.asarray(cvec)
        return {n: r.block.unpack_from(cvec)
                for n, r in self.constraints.items()}

    # -- structure -----------------------------------------------------------------
    @property
    def structure(self):
        if self._structure is None:
            self._structure = Structure(self)
        return self._structure

    def _var_index(self, fun, arg, rows):
        """[rows, core] global decision indices of a function argument."""
        ref = fun['args'][arg]
        core = fun['spec'].core_size(arg)
        if ref[0] == 'param':
            base = self.decision[self.structure.var_names[ref[1]]].offset
            ind = base + np.arange(core, dtype=np.int64)
            return np.broadcast_to(ind, (rows, core))
        assert ref[0] == 'var', (fun['name'], arg, ref)
        base = self.decision[self.structure.var_names[ref[1]]].offset
        k = np.arange(rows, dtype=np.int64)[:, None] + ref[2]
        return base + k * core + np.arange(core, dtype=np.int64)

    def _rows(self, fun):
        reg = self.objectives.get(fun['name']) or self.constraints[fun['name']]
        return reg.rows

    def constr_jac_ind(self):
        """(row, col) index arrays of the Jacobian COO values."""
        if 'jac' not in self._index_cache:
            st = self.structure
            rows, cols = [], []
            for blk in st.jac_blocks:
                fun = st.funs[blk['fun']]
                M = self._rows(fun)
                reg = self.constraints[fun['name']]
                wflat = np.array([e.index[0] for e in blk['entries']])
                oflat = np.array([e.index[1] for e in blk['entries']])
                k = np.arange(M, dtype=np.int64)[:, None]
                rows.append((reg.block.offset + k * fun['out_core']
                             + oflat).ravel())
                cols.append(self._var_index(fun, blk['wrt'], M)[:, wflat]
                            .ravel())
            self._index_cache['jac'] = (
                np.concatenate(rows) if rows else np.zeros(0, np.int64),
                np.concatenate(cols) if cols else np.zeros(0, np.int64))
        return self._index_cache['jac']

    def lag_hess_ind(self):
        """(row, col) index arrays of the Lagrangian-Hessian COO values
        (lower triangle: row >= col)."""
        if 'hess' not in self._index_cache:
            st = self.structure
            rows, cols = [], []
            for blk in st.hess_blocks:
                fun = st.funs[blk['fun']]
                M = self._rows(fun)
                w0, w1 = blk['pair']
                f0 = np.array([e.index[0] for e in blk['entries']])
                f1 = np.array([e.index[1] for e in blk['entries']])
                i0 = self._var_index(fun, w0, M)[:, f0].ravel()
                i1 = self._var_index(fun, w1, M)[:, f1].ravel()
                rows.append(np.maximum(i0, i1))
                cols.append(np.minimum(i0, i1))
            self._index_cache['hess'] = (
                np.concatenate(rows) if rows else np.zeros(0, np.int64),
                np.concatenate(cols) if cols else np.zeros(0, np.int64))
        return self._index_cache['hess']

    @property
    def nnzjac(self):
        st = self.structure
        return sum(self._rows(st.funs[b['fun']]) * b['c']
                   for b in st.jac_blocks)

    @property
    def nnzhess(self):
        st = self.structure
        return sum(self._rows(st.funs[b['fun']]) * b['c']
                   for b in st.hess_blocks)

    # -- GPU evaluation ---------------------------------------------------------------
    @property
    def backend(self):
        """The CUDA evaluator of this problem (compiled on first use)."""
        if self._backend is None:
            from . import backend
            self._backend = backend.ProblemBackend(self)
        return self._backend

    def obj(self, dvec):
        """Objective value (IPOPT eval_f)."""
        return self.backend.eval_f(dvec)

    def obj_grad(self, dvec):
        """Dense objective gradient (IPOPT eval_grad_f)."""
        return self.backend.eval_grad_f(dvec)

    def constr(self, dvec):
        """Constraint vector (IPOPT eval_g)."""
        return self.backend.eval_g(dvec)

    def constr_jac_val(self, dvec):
        """Constraint Jacobian COO values (IPOPT eval_jac_g)."""
        return self.backend.eval_jac_values(dvec)

    def lag_hess_val(self, dvec, obj_mult, constr_mult):
        """Lagrangian Hessian COO values (IPOPT eval_h)."""
        return self.backend.eval_hess_values(dvec, obj_mult, constr_mult)

    @contextlib.contextmanager
    def ipopt(self, dec_bounds, constr_bounds):
        """Context manager yielding the NLP solver (attas_sp_ml.py:153)."""
        from . import nlp
        solver = nlp.make_solver(self, dec_bounds, constr_bounds)
        try:
            yield solver
        finally:
            solver.close()
