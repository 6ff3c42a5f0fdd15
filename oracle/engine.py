"""CPU ORACLE ENGINE -- TEST INFRASTRUCTURE ONLY, NEVER A PRODUCT PATH.

A clean-room, CPU-only stand-in for the two third-party packages the reference
builds on but does not ship (``ceacoest.modelling.symoptim`` + ``sym2num`` for
the model side, ``ceacoest.optim`` for the problem side; call sites
/root/reference/fem.py:6-62, /root/reference/symfem.py:8-48,
/root/reference/attas_sp_ml.py:85-87,153-159).  It differentiates the model
expressions with sympy, evaluates one lambdified NumPy broadcast expression per
structural nonzero (the way sym2num-generated code does) and assembles the
dense gradient and the COO Jacobian / Lagrangian Hessian.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package (``colloc_fem_code_b200``) never does.

PARITY STATUS: the reference has no tests, golden vectors or fixtures, and the
packages that define its COO order are absent, so the reference pins nothing
("parity unpinned" by the reference's own tests).  The pins used instead are
fixtures generated in the build container by running the reference's own
``symfem.py`` / ``fem.py`` UNCHANGED on top of this engine
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``).

The COO ordering contract (stated once, shared with DESIGN.md):

* Jacobian: constraints in registration order; inside a constraint, decision
  arguments in signature order (cf. /root/reference/adfem.py:198-203); inside a
  (function, wrt) block the sample index is slowest (cf. adfem.py:303-328) and
  the structural nonzeros of one sample are in C order of
  ``wrt.core_shape + out.core_shape``.
* Hessian of the Lagrangian: objectives first, then constraints, each in
  registration order; inside a function the argument pairs of
  ``itertools.combinations_with_replacement`` (adfem.py:206-209); inside a block
  sample-major, core nonzeros in C order of ``w0.core + w1.core + out.core``;
  same-variable blocks keep ``flat(w0) >= flat(w1)`` (lower triangle WITH the
  diagonal, the intent of adfem.py:41-57 and what IPOPT needs); every entry is
  oriented ``row >= col`` in global decision indices.
* Parameter (non per-sample) variables of a per-sample function are broadcast,
  i.e. their entries repeat once per sample (adfem.py:316-318); IPOPT sums
  duplicates.
* Objective gradient: dense; parameter entries are summed over the samples
  (adfem.py:119).
"""

import collections
import inspect
import itertools

import numpy as np
import sympy


# ----------------------------------------------------------------------------
# Model side (stand-in for ceacoest.modelling.symoptim.Model + sym2num)
# ----------------------------------------------------------------------------

class SymbolTable(dict):
    """``model.variables``: assigning (nested lists of) names creates symbols.

    Mirrors how /root/reference/symfem.py:24-43 fills ``self.variables``.
    """

    def __setitem__(self, name, spec):
        if name != 'self':
            spec = _symbolize(spec)
        super().__setitem__(name, spec)


def _symbolize(spec):
    if isinstance(spec, str):
        return sympy.Symbol(spec, real=True)
    if isinstance(spec, sympy.Basic):
        return spec
    arr = np.array(spec, dtype=object)
    out = np.empty(arr.shape, dtype=object)
    for ind in np.ndindex(*arr.shape):
        elem = arr[ind]
        out[ind] = _symbolize(elem) if isinstance(elem, str) else elem
    return out


class SymFunction:
    """One symbolic model function with its sparse first/second derivatives."""

    def __init__(self, model, name):
        self.name = name
        self.__name__ = name
        method = getattr(model, name)
        self.args = list(inspect.signature(method).parameters)
        self.__signature__ = inspect.signature(method)
        v = model.variables
        self.arg_syms = {a: np.asarray(v[a], dtype=object) for a in self.args}
        out = method(*(v[a] for a in self.args))
        self.out = np.asarray(out, dtype=object)
        self.core_shape = self.out.shape
        self.wrt = [a for a in self.args if a in model.decision]

        # symbol -> (argument name, flat index in the argument's core)
        owner = {}
        for a in self.wrt:
            for flat, s in enumerate(self.arg_syms[a].ravel()):
                owner[s] = (a, flat)

        jac = {a: [] for a in self.wrt}
        pairs = list(itertools.combinations_with_replacement(self.wrt, 2))
        hess = {p: [] for p in pairs}
        pos = {a: i for i, a in enumerate(self.wrt)}
        for oflat, expr in enumerate(self.out.ravel()):
            expr = sympy.sympify(expr)
            for s1 in sorted(expr.free_symbols, key=str):
                if s1 not in owner:
                    continue
                d1 = sympy.diff(expr, s1)
                if d1 == 0:
                    continue
                a1, i1 = owner[s1]
                jac[a1].append((i1, oflat, d1))
                for s2 in sorted(d1.free_symbols, key=str):
                    if s2 not in owner:
                        continue
                    a2, i2 = owner[s2]
                    if pos[a2] < pos[a1]:
                        continue        # visited from the other side
                    if a1 == a2 and i1 < i2:
                        continue        # lower triangle incl. diagonal
                    d2 = sympy.diff(d1, s2)
                    if d2 == 0:
                        continue
                    hess[(a1, a2)].append((i1, i2, oflat, d2))
        self.jac = {a: sorted(e, key=lambda t: t[:2]) for a, e in jac.items()}
        self.hess = {p: sorted(e, key=lambda t: t[:3])
                     for p, e in hess.items() if e}
        self._lambdas = {}

    def __repr__(self):
        return f"<oracle SymFunction '{self.name}'>"

    # -- numeric evaluation (one NumPy expression per structural nonzero) ----
    def _lam(self, key, exprs):
        try:
            return self._lambdas[key]
        except KeyError:
            flat_args = [list(self.arg_syms[a].ravel()) for a in self.args]
            f = sympy.lambdify(flat_args, list(exprs), modules='numpy',
                               cse=False)
            self._lambdas[key] = f
            return f

    def _split(self, values):
        """Core-flattened, sample-leading argument arrays for lambdify."""
        out = []
        for a in self.args:
            core = self.arg_syms[a].shape
            val = np.asarray(values[a], dtype=float)
            ext = val.shape[:val.ndim - len(core)]
            flat = val.reshape(ext + (-1,))
            out.append([flat[..., i] for i in range(flat.shape[-1])])
        return out

    def value(self, values, ext):
        f = self._lam('val', self.out.ravel())
        cols = f(*self._split(values))
        res = np.empty(ext + (self.out.size,))
        for i, c in enumerate(cols):
            res[..., i] = c
        return res.reshape(ext + self.core_shape)

    def block_values(self, key, entries, values, ext):
        """[ext..., nnz_core] array of one derivative block."""
        f = self._lam(key, [e[-1] for e in entries])
        cols = f(*self._split(values))
        res = np.empty(ext + (len(entries),))
        for i, c in enumerate(cols):
            res[..., i] = c
        return res


class Model:
    """Stand-in for ``ceacoest.modelling.symoptim.Model``."""

    generated_name = None

    def __init__(self):
        self.variables = SymbolTable()
        self.variables['self'] = {}
        self.decision = set()
        self.constraints = []
        self.objectives = []

    def add_constraint(self, name):
        if name not in self.constraints:
            self.constraints.append(name)

    def add_objective(self, name):
        if name not in self.objectives:
            self.objectives.append(name)

    def compile_class(self):
        symmodel = self
        assigns = dict(getattr(self, 'generate_assignments', {}))
        funcs = {n: SymFunction(self, n)
                 for n in self.objectives + self.constraints}

        class CompiledOracleModel:
            symbolic = symmodel
            functions = funcs
            decision = frozenset(symmodel.decision)

            def __init__(self):
                for k, val in assigns.items():
                    setattr(self, k, val)
                for n, f in funcs.items():
                    setattr(self, n, f)

        name = self.generated_name or ('Generated' + type(self).__name__)
        CompiledOracleModel.__name__ = name
        return CompiledOracleModel

    def print_code(self):
        raise NotImplementedError(
            'the oracle does not emit source; use compile_class()')


# ----------------------------------------------------------------------------
# Problem side (stand-in for ceacoest.optim)
# ----------------------------------------------------------------------------

class Decision:
    """A block of the decision (or constraint) vector: shape + offset."""

    def __init__(self, shape, offset):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.offset = int(offset)
        self.size = int(np.prod(self.shape, dtype=np.int64))

    def unpack_from(self, vec):
        return vec[self.offset:self.offset + self.size].reshape(self.shape)


class _Registered:
    def __init__(self, fun, shape, offset, problem):
        self.fun = fun
        self.spec = Decision(shape, offset)
        ncore = len(fun.core_shape)
        shape = self.spec.shape
        self.ext = shape[:len(shape) - ncore]
        assert shape[len(self.ext):] == tuple(fun.core_shape), \
            (fun.name, shape, fun.core_shape)
        self.M = int(np.prod(self.ext, dtype=np.int64))
        self.out_core = int(np.prod(fun.core_shape, dtype=np.int64))
        self.problem = problem

    def var_index(self, arg):
        """Global decision index of every (sample, core) element of ``arg``.

        Returns an ``[M, core]`` integer array; parameters broadcast over M.
        """
        p = self.problem
        spec = p.dependent.get(arg) or p.decision[arg]
        core = int(np.prod(self.fun.arg_syms[arg].shape, dtype=np.int64))
        ind = spec.offset + np.arange(spec.size, dtype=np.int64)
        ind = ind.reshape(-1, core)
        if len(ind) == 1:
            ind = np.broadcast_to(ind, (self.M, core))
        assert ind.shape == (self.M, core), (self.fun.name, arg, ind.shape)
        return ind


class Problem:
    """Stand-in for ``ceacoest.optim.Problem`` with NumPy evaluation."""

    def __init__(self):
        self.decision = collections.OrderedDict()
        self.dependent = collections.OrderedDict()
        self.constraints = collections.OrderedDict()
        self.objectives = collections.OrderedDict()
        self.ndec = 0
        self.ncons = 0

    # -- registration ---------------------------------------------------------
    def add_decision(self, name, shape):
        spec = Decision(shape, self.ndec)
        self.decision[name] = spec
        self.ndec += spec.size
        return spec

    def add_dependent_variable(self, name, spec):
        self.dependent[name] = spec

    def add_objective(self, fun, shape):
        self.objectives[fun.__name__] = _Registered(fun, shape, 0, self)

    def add_constraint(self, fun, shape):
        reg = _Registered(fun, shape, self.ncons, self)
        self.constraints[fun.__name__] = reg
        self.ncons += reg.spec.size

    # -- packing --------------------------------------------------------------
    def variables(self, dvec):
        dvec = np.asarray(dvec)
        out = {n: s.unpack_from(dvec) for n, s in self.decision.items()}
        out.update({n: s.unpack_from(dvec) for n, s in self.dependent.items()})
        return out

    def unpack_constraints(self, cvec):
        cvec = np.asarray(cvec)
        return {n: r.spec.unpack_from(cvec)
                for n, r in self.constraints.items()}

    # -- evaluation -----------------------------------------------------------
    def obj(self, dvec):
        var = self.variables(np.asarray(dvec, dtype=float))
        total = 0.0
        for reg in self.objectives.values():
            total += reg.fun.value(var, reg.ext).sum()
        return float(total)

    def obj_grad(self, dvec):
        var = self.variables(np.asarray(dvec, dtype=float))
        grad = np.zeros(self.ndec)
        for reg in self.objectives.values():
            for wrt, entries in reg.fun.jac.items():
                if not entries:
                    continue
                vals = reg.fun.block_values(('jac', wrt), entries, var,
                                            reg.ext).reshape(reg.M, -1)
                vind = reg.var_index(wrt)
                cols = vind[:, [e[0] for e in entries]]
                np.add.at(grad, cols.ravel(), vals.ravel())
        return grad

    def constr(self, dvec):
        var = self.variables(np.asarray(dvec, dtype=float))
        g = np.empty(self.ncons)
        for reg in self.constraints.values():
            reg.spec.unpack_from(g)[...] = reg.fun.value(var, reg.ext)
        return g

    def _jac_blocks(self):
        for reg in self.constraints.values():
            for wrt in reg.fun.wrt:
                entries = reg.fun.jac[wrt]
                if entries:
                    yield reg, wrt, entries

    def constr_jac_ind(self):
        rows, cols = [], []
        for reg, wrt, entries in self._jac_blocks():
            k = np.arange(reg.M, dtype=np.int64)[:, None]
            oflat = np.array([e[1] for e in entries], dtype=np.int64)
            rows.append((reg.spec.offset + k * reg.out_core + oflat).ravel())
            vind = reg.var_index(wrt)
            cols.append(vind[:, [e[0] for e in entries]].ravel())
        return np.concatenate(rows), np.concatenate(cols)

    def constr_jac_val(self, dvec):
        var = self.variables(np.asarray(dvec, dtype=float))
        vals = []
        for reg, wrt, entries in self._jac_blocks():
            vals.append(reg.fun.block_values(('jac', wrt), entries, var,
                                             reg.ext).ravel())
        return np.concatenate(vals)

    def _hess_blocks(self):
        regs = itertools.chain(
            ((r, True) for r in self.objectives.values()),
            ((r, False) for r in self.constraints.values()))
        for reg, is_obj in regs:
            for pair, entries in reg.fun.hess.items():
                yield reg, is_obj, pair, entries

    def lag_hess_ind(self):
        rows, cols = [], []
        for reg, is_obj, (w0, w1), entries in self._hess_blocks():
            i0 = reg.var_index(w0)[:, [e[0] for e in entries]].ravel()
            i1 = reg.var_index(w1)[:, [e[1] for e in entries]].ravel()
            rows.append(np.maximum(i0, i1))
            cols.append(np.minimum(i0, i1))
        return np.concatenate(rows), np.concatenate(cols)

    def lag_hess_val(self, dvec, obj_mult, constr_mult):
        var = self.variables(np.asarray(dvec, dtype=float))
        constr_mult = np.asarray(constr_mult, dtype=float)
        vals = []
        for reg, is_obj, pair, entries in self._hess_blocks():
            d2 = reg.fun.block_values(('hess', pair), entries, var,
                                      reg.ext).reshape(reg.M, -1)
            if is_obj:
                mult = obj_mult
            else:
                lam = reg.spec.unpack_from(constr_mult)
                lam = lam.reshape(reg.M, reg.out_core)
                mult = lam[:, [e[2] for e in entries]]
            vals.append((d2 * mult).ravel())
        return np.concatenate(vals)

    @property
    def nnzjac(self):
        return sum(reg.M * len(e) for reg, _, e in self._jac_blocks())

    @property
    def nnzhess(self):
        return sum(reg.M * len(e) for reg, _, _, e in self._hess_blocks())

    def ipopt(self, dec_bounds, constr_bounds):
        raise NotImplementedError('the oracle has no solver binding')
