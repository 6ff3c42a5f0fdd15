"""Symbolic model front-end: the ``ceacoest.modelling.symoptim.Model`` boundary.

The reference's models subclass ``symoptim.Model`` (/root/reference/symfem.py:8,
11-48): they fill ``self.variables`` with (nested lists of) symbol names, add
names to the ``self.decision`` set, register functions by name with
``add_constraint`` / ``add_objective`` and are turned into a numeric model with
``compile_class()()`` (/root/reference/attas_sp_ml.py:85-86) or
``print_code()`` (/root/reference/hfb320_sqrt_zoh.py:41-48,
/root/reference/mc_blackbox_cfem.py:81-95).  This module keeps that surface.

What differs from the reference stack: nothing here generates NumPy code.  The
model functions are differentiated with sympy ONCE, reduced to their structural
nonzeros, and every expression is printed as a C expression over canonical
identifiers (``v_<arg>_<flat>``).  The resulting :class:`ModelSpec` is plain
data (JSON-serialisable) and is what ``codegen.py`` turns into the fused FP64
CUDA kernels; there is no CPU evaluation path in this package.

Derivative/ordering conventions (shared with DESIGN.md and ``oracle/engine.py``):
first derivatives with respect to every decision argument in signature order
(cf. /root/reference/adfem.py:198-203); second derivatives for the argument
pairs of ``combinations_with_replacement`` (adfem.py:206-209); structural
nonzeros of one sample in C order of ``wrt.core + out.core``; same-variable
Hessian blocks keep the lower triangle including the diagonal.
"""

import inspect
import itertools
import json

import numpy as np
import sympy
from sympy.printing.c import C99CodePrinter


# ----------------------------------------------------------------------------
# symbol table
# ----------------------------------------------------------------------------

class VariableTable(dict):
    """``model.variables``: names in, sympy symbols out.

    ``v['A'] = [['A0_0', 'A0_1'], ...]`` stores an object ndarray of symbols,
    ``v['dt'] = 'dt'`` a scalar symbol (symfem.py:25-43,187).  The table also
    carries a ``'self'`` entry, which the reference filters out explicitly
    (symfem.py:38).
    """

    def __init__(self):
        super().__init__()
        dict.__setitem__(self, 'self', {})

    def __setitem__(self, name, spec):
        dict.__setitem__(self, name, self._convert(spec))

    @classmethod
    def _convert(cls, spec):
        if isinstance(spec, str):
            return sympy.Symbol(spec, real=True)
        if isinstance(spec, (sympy.Basic, np.ndarray)):
            return spec
        arr = np.empty(np.shape(spec), dtype=object)
        src = np.array(spec, dtype=object)
        for ind in np.ndindex(*arr.shape):
            arr[ind] = cls._convert(src[ind])
        return arr


# ----------------------------------------------------------------------------
# C printing
# ----------------------------------------------------------------------------

class _CExpr(C99CodePrinter):
    """FP64 C expressions; small integer powers become multiplications."""

    def __init__(self, names):
        super().__init__()
        self._names = names

    def _print_Symbol(self, sym):
        return self._names[sym]

    def _print_Pow(self, expr):
        base, exp = expr.as_base_exp()
        if exp.is_Integer and 1 <= abs(int(exp)) <= 4:
            b = self._print(base)
            if not (base.is_Symbol or base.is_Number):
                b = f'({b})'
            prod = '*'.join([b] * abs(int(exp)))
            return prod if exp > 0 else f'(1.0/({prod}))'
        return super()._print_Pow(expr)

    def _print_Integer(self, expr):
        return f'{int(expr)}.0'

    def _print_Rational(self, expr):
        return f'({int(expr.p)}.0/{int(expr.q)}.0)'


# ----------------------------------------------------------------------------
# specs (plain data)
# ----------------------------------------------------------------------------

class ExprSpec:
    """One structural nonzero: its indices, C code and symbol dependencies."""

    __slots__ = ('index', 'code', 'deps', 'terms')

    def __init__(self, index, code, deps, terms=None):
        self.index = tuple(int(i) for i in index)
        self.code = code
        self.deps = sorted((str(a), int(i)) for a, i in deps)
        # additive terms [(code, deps)]: lets the generator hoist the parts of
        # an objective that do not depend on the sample out of the reduction
        self.terms = None if terms is None else [
            (c, sorted((str(a), int(i)) for a, i in d)) for c, d in terms]

    def to_json(self):
        out = [list(self.index), self.code, [list(d) for d in self.deps]]
        if self.terms is not None:
            out.append([[c, [list(x) for x in d]] for c, d in self.terms])
        return out

    @classmethod
    def from_json(cls, data):
        terms = None
        if len(data) > 3:
            terms = [(c, [tuple(x) for x in d]) for c, d in data[3]]
        return cls(data[0], data[1], [tuple(d) for d in data[2]], terms)


class FunctionSpec:
    """A model function reduced to values + sparse first/second derivatives.

    ``args``      ordered argument names (the call signature)
    ``core``      ``{arg: core shape}``
    ``out_core``  core shape of the output
    ``wrt``       decision arguments, signature order
    ``values``    ``ExprSpec`` per output element, index ``(out_flat,)``
    ``jac``       ``{wrt: [ExprSpec index (wrt_flat, out_flat)]}``
    ``hess``      ``{(w0, w1): [ExprSpec index (i0, i1, out_flat)]}``
    """

    def __init__(self, name, args, core, out_core, wrt, values, jac, hess):
        self.name = name
        self.args = list(args)
        self.core = {a: tuple(s) for a, s in core.items()}
        self.out_core = tuple(out_core)
        self.wrt = list(wrt)
        self.values = values
        self.jac = jac
        self.hess = hess

    @property
    def out_size(self):
        return int(np.prod(self.out_core, dtype=np.int64))

    def core_size(self, arg):
        return int(np.prod(self.core[arg], dtype=np.int64))

    def to_json(self):
        return {
            'name': self.name, 'args': self.args,
            'core': {a: list(s) for a, s in self.core.items()},
            'out_core': list(self.out_core), 'wrt': self.wrt,
            'values': [e.to_json() for e in self.values],
            'jac': [[w, [e.to_json() for e in es]]
                    for w, es in self.jac.items()],
            'hess': [[list(p), [e.to_json() for e in es]]
                     for p, es in self.hess.items()],
        }

    @classmethod
    def from_json(cls, d):
        ex = ExprSpec.from_json
        return cls(d['name'], d['args'], d['core'], d['out_core'], d['wrt'],
                   [ex(e) for e in d['values']],
                   {w: [ex(e) for e in es] for w, es in d['jac']},
                   {tuple(p): [ex(e) for e in es] for p, es in d['hess']})


class ModelSpec:
    """Everything ``compile_class`` learned about a symbolic model."""

    def __init__(self, name, assignments, decision, objectives, constraints,
                 functions):
        self.name = name
        self.assignments = dict(assignments)
        self.decision = sorted(decision)
        self.objectives = list(objectives)
        self.constraints = list(constraints)
        self.functions = functions          # {name: FunctionSpec}

    def to_json(self):
        return {
            'name': self.name, 'assignments': self.assignments,
            'decision': self.decision, 'objectives': self.objectives,
            'constraints': self.constraints,
            'functions': [f.to_json() for f in self.functions.values()],
        }

    @classmethod
    def from_json(cls, d):
        funs = [FunctionSpec.from_json(f) for f in d['functions']]
        return cls(d['name'], d['assignments'], d['decision'],
                   d['objectives'], d['constraints'],
                   {f.name: f for f in funs})

    def dumps(self):
        return json.dumps(self.to_json(), sort_keys=True)


def c_ident(arg, flat):
    """Canonical C identifier of element ``flat`` of argument ``arg``."""
    return f'v_{arg}_{flat}'


def analyse_function(model, name):
    """Differentiate one registered model function into a FunctionSpec."""
    method = getattr(model, name)
    args = list(inspect.signature(method).parameters)
    table = model.variables
    syms = {a: np.asarray(table[a], dtype=object) for a in args}
    out = np.asarray(method(*(table[a] for a in args)), dtype=object)

    owner, names = {}, {}
    for a in args:
        for flat, s in enumerate(syms[a].ravel()):
            owner[s] = (a, flat)
            names[s] = c_ident(a, flat)
    printer = _CExpr(names)
    wrt = [a for a in args if a in model.decision]
    rank = {a: i for i, a in enumerate(wrt)}

    split_terms = name in model.objectives

    def spec(index, expr):
        expr = sympy.sympify(expr)
        unknown = [s for s in expr.free_symbols if s not in owner]
        if unknown:
            raise ValueError(f'{name}: expression uses symbols {unknown} '
                             'that are not arguments of the function')
        terms = None
        if split_terms:
            terms = [(printer.doprint(t), [owner[s] for s in t.free_symbols])
                     for t in sympy.Add.make_args(expr)]
        return ExprSpec(index, printer.doprint(expr),
                        [owner[s] for s in expr.free_symbols], terms)

    values, jac, hess = [], {a: {} for a in wrt}, {}
    for oflat, expr in enumerate(out.ravel()):
        expr = sympy.sympify(expr)
        values.append(spec((oflat,), expr))
        dsyms = [s for s in expr.free_symbols
                 if s in owner and owner[s][0] in rank]
        for s1 in dsyms:
            a1, i1 = owner[s1]
            d1 = expr.diff(s1)
            if d1 == 0:
                continue
            jac[a1][(i1, oflat)] = d1
            for s2 in d1.free_symbols:
                if s2 not in owner or owner[s2][0] not in rank:
                    continue
                a2, i2 = owner[s2]
                if rank[a2] < rank[a1] or (a1 == a2 and i2 > i1):
                    continue
                d2 = d1.diff(s2)
                if d2 != 0:
                    hess.setdefault((a1, a2), {})[(i1, i2, oflat)] = d2

    jac_spec = {a: [spec(k, e) for k, e in sorted(jac[a].items())]
                for a in wrt if jac[a]}
    pairs = itertools.combinations_with_replacement(wrt, 2)
    hess_spec = {p: [spec(k, e) for k, e in sorted(hess[p].items())]
                 for p in pairs if p in hess}
    return FunctionSpec(name, args, {a: syms[a].shape for a in args},
                        out.shape, wrt, values, jac_spec, hess_spec)


# ----------------------------------------------------------------------------
# the compiled (numeric-side) model object handed to the Problem classes
# ----------------------------------------------------------------------------

class ModelFunction:
    """Handle of one compiled model function (``model.dynamics`` etc.).

    Problems receive these through ``add_constraint(model.dynamics, shape)``
    (/root/reference/fem.py:55-57); they carry the name/signature the glue
    layer binds arguments by (cf. /root/reference/adfem.py:25-31).  Evaluation
    happens inside the fused CUDA kernels of the owning Problem, so calling a
    handle directly is an error rather than a silent CPU evaluation.
    """

    def __init__(self, model, spec):
        self.model = model
        self.spec = spec
        self.__name__ = spec.name
        params = [inspect.Parameter(a, inspect.Parameter.POSITIONAL_OR_KEYWORD)
                  for a in spec.args]
        self.__signature__ = inspect.Signature(params)

    def __repr__(self):
        return f"<colloc_fem_code_b200 model function '{self.__name__}'>"

    def __call__(self, *args, **kwargs):
        raise RuntimeError(
            f"'{self.__name__}' is evaluated on the GPU by the Problem it is "
            "registered with (obj/constr/constr_jac_val/lag_hess_val); there "
            "is no stand-alone CPU evaluation in this package")

    # -- structural half of the per-function operator interface --------------
    # (the in-tree statement of what the glue asks of a model function is
    # /root/reference/adfem.py:20-120: jac_nnz / jac_ind / jac_val,
    # hess_nnz / hess_ind / hess_val, grad).  Indices are variable-local, the
    # sample index is slowest (adfem.py:303-328); unlike adfem's dense blocks
    # only structural nonzeros are listed, same-variable Hessian blocks keep
    # the lower triangle including the diagonal.
    def _ext(self, shape, core):
        shape = tuple(np.atleast_1d(shape)) if not isinstance(shape, tuple) \
            else shape
        ncore = len(core)
        ext = shape[:len(shape) - ncore]
        if tuple(shape[len(ext):]) != tuple(core):
            raise ValueError(f'{self.__name__}: shape {shape} does not end '
                             f'with the core shape {tuple(core)}')
        return int(np.prod(ext, dtype=np.int64)), len(ext) > 0

    def _local(self, arg, dec_shapes, flats, M):
        """[M, nnz] variable-local flat indices of core entries ``flats``."""
        spec = self.spec
        core = int(np.prod(spec.core[arg], dtype=np.int64))
        rows, per_sample = self._ext(tuple(dec_shapes[arg]), spec.core[arg])
        flats = np.asarray(flats, dtype=np.int64)
        if not per_sample:
            return np.broadcast_to(flats, (M, len(flats)))
        if rows != M:
            raise ValueError(f'{self.__name__}: {arg} has {rows} rows, the '
                             f'output has {M}')
        return np.arange(M, dtype=np.int64)[:, None] * core + flats

    def jac_nnz(self, dec_shapes, out_shape):
        M, _ = self._ext(tuple(np.atleast_1d(out_shape)), self.spec.out_core)
        return M * sum(len(e) for e in self.spec.jac.values())

    def jac_ind(self, dec_shapes, out_shape):
        """``{(wrt,): int array (2, nnz)}`` rows ``[wrt index, out index]``."""
        import collections
        spec = self.spec
        M, _ = self._ext(tuple(np.atleast_1d(out_shape)), spec.out_core)
        out = collections.OrderedDict()
        k = np.arange(M, dtype=np.int64)[:, None]
        for wrt, entries in spec.jac.items():
            w = self._local(wrt, dec_shapes, [e.index[0] for e in entries], M)
            o = k * spec.out_size + np.array([e.index[1] for e in entries])
            out[(wrt,)] = np.array([w.ravel(), o.ravel()])
        return out

    def hess_nnz(self, dec_shapes, out_shape):
        M, _ = self._ext(tuple(np.atleast_1d(out_shape)), self.spec.out_core)
        return M * sum(len(e) for e in self.spec.hess.values())

    def hess_ind(self, dec_shapes, out_shape):
        """``{(w0, w1): int array (3, nnz)}`` rows ``[w0, w1, out index]``."""
        import collections
        spec = self.spec
        M, _ = self._ext(tuple(np.atleast_1d(out_shape)), spec.out_core)
        out = collections.OrderedDict()
        k = np.arange(M, dtype=np.int64)[:, None]
        for (w0, w1), entries in spec.hess.items():
            i0 = self._local(w0, dec_shapes, [e.index[0] for e in entries], M)
            i1 = self._local(w1, dec_shapes, [e.index[1] for e in entries], M)
            o = k * spec.out_size + np.array([e.index[2] for e in entries])
            out[(w0, w1)] = np.array([i0.ravel(), i1.ravel(), o.ravel()])
        return out

    def _gpu_only(self, what):
        raise RuntimeError(
            f"{self.__name__}.{what}: derivative VALUES are produced by the "
            "fused CUDA kernels of the Problem this function is registered "
            "with (constr_jac_val / lag_hess_val / obj_grad)")

    def jac_val(self, *args, **kwargs):
        self._gpu_only('jac_val')

    def hess_val(self, *args, **kwargs):
        self._gpu_only('hess_val')

    def grad(self, *args, **kwargs):
        self._gpu_only('grad')


class CompiledModel:
    """Base of the classes returned by ``Model.compile_class()``."""

    spec = None     # ModelSpec, set on the generated subclass

    def __init__(self):
        for key, val in self.spec.assignments.items():
            setattr(self, key, val)
        for name, fspec in self.spec.functions.items():
            setattr(self, name, ModelFunction(self, fspec))

    @property
    def decision(self):
        return set(self.spec.decision)


def compiled_class(spec):
    """Class object for a ModelSpec (used by compile_class and print_code)."""
    return type(spec.name, (CompiledModel,), {'spec': spec})


def compiled_class_from_json(text):
    return compiled_class(ModelSpec.from_json(json.loads(text)))


class Model:
    """Drop-in for ``ceacoest.modelling.symoptim.Model`` (symfem.py:11-48)."""

    generated_name = None
    """Name of the generated class (mc_blackbox_cfem.py:26)."""

    def __init__(self):
        self.variables = VariableTable()
        self.decision = set()
        self.constraints = []
        self.objectives = []

    def add_constraint(self, name):
        if name not in self.constraints:
            self.constraints.append(name)

    def add_objective(self, name):
        if name not in self.objectives:
            self.objectives.append(name)

    def model_spec(self):
        name = self.generated_name or f'Generated{type(self).__name__}'
        assigns = {k: (int(v) if isinstance(v, (int, np.integer)) else v)
                   for k, v in getattr(self, 'generate_assignments',
                                       {}).items()}
        funs = {n: analyse_function(self, n)
                for n in self.objectives + self.constraints}
        return ModelSpec(name, assigns, self.decision, self.objectives,
                         self.constraints, funs)

    def compile_class(self):
        """Class whose instances are the compiled model (attas_sp_ml.py:86)."""
        return compiled_class(self.model_spec())

    def print_code(self):
        """Source of a module defining the compiled class, for the reference's
        write-and-import cache (mc_blackbox_cfem.py:81-95)."""
        spec = self.model_spec()
        return (
            '"""Generated by colloc_fem_code_b200.symoptim -- do not edit."""\n'
            'from colloc_fem_code_b200.symoptim import '
            'compiled_class_from_json\n\n'
            f'_SPEC = {spec.dumps()!r}\n\n'
            f'{spec.name} = compiled_class_from_json(_SPEC)\n')
