"""Make the reference's scripts and modules resolve to this package.

The reference's ``symfem.py`` / ``fem.py`` import ``ceacoest.modelling.symoptim``
and ``ceacoest.optim`` (symfem.py:8, fem.py:6) and the scripts import
``sym2num.model`` (attas_sp_ml.py:10).  :func:`install` registers modules of
those names in ``sys.modules`` that forward to :mod:`colloc_fem_code_b200.symoptim`
and :mod:`colloc_fem_code_b200.optim`, so the reference's own files run unchanged
on the CUDA path; when the reference tree is not importable the names ``symfem``
and ``fem`` are bound to the mirrors :mod:`.models` / :mod:`.problems`.

Stale class names used by ``hfb320_sqrt_zoh.py`` and ``blackbox_innov_bal.py``
(classes that no longer exist in the reference's HEAD)
are defined on the mirrors from their nearest HEAD equivalents; their
``variables()`` also answers to the older parametrisation's variable names
(``L``, ``e``, ``W_diag``, ``isRp_tril``, ``Qc``), and the generated-model
modules those scripts import are produced on demand.
"""

import importlib.abc
import importlib.machinery
import re
import sys
import types

import numpy as np

from . import models, optim, problems, symoptim


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__dict__['__cfem_compat__'] = True
    return mod


# -----------------------------------------------------------------------------
# variables of the older parametrisation (hfb320_sqrt_zoh.py:113-125,
# blackbox_innov_bal.py:69-73)
# -----------------------------------------------------------------------------

class DerivedView:
    """Array-like stand-in for a variable of the older parametrisation that
    HEAD replaced by a function of it (``W_diag = sW_diag**2``,
    ``isRp = sRp**-1``, ``Qc = sQc sQc'``).

    Reads give the value derived from the HEAD variable when the view was
    made.  Writes (``view[idx] = value``) update that value and, when it maps
    back to a valid HEAD value, write the HEAD variable in place; content that
    has no image (the -inf / +inf patterns of a bounds vector, a scale
    vector's constants) is kept in the view only -- a bound or scale on the
    old variable does not transfer to the new one -- and reported once
    through ``warnings``.
    """

    def __init__(self, name, value, push):
        self.name = name
        self._val = np.array(value, dtype=float)
        self._push = push
        self._warned = False

    shape = property(lambda self: self._val.shape)
    size = property(lambda self: self._val.size)
    ndim = property(lambda self: self._val.ndim)
    dtype = property(lambda self: self._val.dtype)

    def __len__(self):
        return len(self._val)

    def __array__(self, dtype=None, copy=None):
        return np.array(self._val, dtype=dtype)

    def __getitem__(self, idx):
        return self._val[idx]

    def __setitem__(self, idx, value):
        self._val[idx] = value
        if not self._push(self._val) and not self._warned:
            import warnings
            self._warned = True
            warnings.warn(
                f'compat: {self.name!r} belongs to an older parametrisation; '
                'this content has no image in the current variables and is '
                'not transferred', stacklevel=2)

    def __repr__(self):
        return f'DerivedView({self.name!r}, {self._val!r})'


def _tril(elem, n):
    m = np.zeros((n, n))
    m[np.tril_indices(n)] = elem
    return m


def _owned(target, last):
    """A HEAD block may be overwritten from the old variable if nothing else
    has been put there: it is still all zero (a fresh guess vector) or holds
    what this view wrote last."""
    return not target.any() or (last is not None
                                and np.array_equal(target, last))


def _squared_view(sw):
    """``W_diag`` over ``sW_diag`` (blackbox_innov_bal.py:72,80,113):
    elementwise, so bounds transfer too (0 <-> 0, +-inf <-> +-inf)."""
    def push(w):
        ok = ~np.isnan(w) & ((w >= 0) | np.isinf(w))
        with np.errstate(invalid='ignore'):
            sw[ok] = np.where(np.isinf(w[ok]), w[ok], np.sqrt(np.abs(w[ok])))
        return bool(ok.all())
    return DerivedView('W_diag', sw * np.abs(sw), push)


def _inverse_tril_view(srp_tril):
    """``isRp_tril`` over ``sRp_tril`` (hfb320_sqrt_zoh.py:118,139;
    blackbox_innov_bal.py:73): ``sRp = isRp**-1`` as triangular matrices."""
    n = models.tril_mat(np.zeros(srp_tril.size)).shape[0]
    tri = np.tril_indices(n)
    state = {'last': None}

    def invert(elem):
        m = _tril(elem, n)
        if not np.all(np.isfinite(m)) or not np.all(np.diag(m) != 0):
            return None
        return np.linalg.inv(m)[tri]

    def push(val):
        new = invert(val)
        if new is None or not _owned(srp_tril, state['last']):
            return False
        srp_tril[...] = new
        state['last'] = new.copy()
        return True
    start = invert(srp_tril)
    return DerivedView('isRp_tril',
                       np.zeros(srp_tril.size) if start is None else start,
                       push)


def _gram_view(sqc_tril):
    """``Qc`` over ``sQc_tril`` (hfb320_sqrt_zoh.py:125,148,169,186):
    ``Qc = sQc sQc'``."""
    n = models.tril_mat(np.zeros(sqc_tril.size)).shape[0]
    tri = np.tril_indices(n)
    state = {'last': None}

    def push(val):
        if not np.all(np.isfinite(val)) or not _owned(sqc_tril, state['last']):
            return False
        try:
            new = np.linalg.cholesky(0.5 * (val + val.T))[tri]
        except np.linalg.LinAlgError:
            return False
        sqc_tril[...] = new
        state['last'] = new.copy()
        return True
    s = _tril(sqc_tril, n)
    start = s @ s.T if np.all(np.isfinite(s)) else np.zeros((n, n))
    return DerivedView('Qc', start, push)


def stale_variable_aliases(var):
    """Add the older parametrisation's names to a ``problem.variables()``
    dict: ``L`` and ``e`` are the SAME views as ``Ln`` and ``en`` (HEAD
    normalises the innovations: ``en = sRp**-1 e``, ``Ln = L sRp``; as initial
    guesses and for plotting the scripts only need the storage), ``W_diag``,
    ``isRp_tril`` and ``Qc`` are :class:`DerivedView` objects."""
    for old, new in (('L', 'Ln'), ('e', 'en')):
        if new in var and old not in var:
            var[old] = var[new]
    if 'sW_diag' in var and 'W_diag' not in var:
        var['W_diag'] = _squared_view(var['sW_diag'])
    if 'sRp_tril' in var and 'isRp_tril' not in var:
        var['isRp_tril'] = _inverse_tril_view(var['sRp_tril'])
    if 'sQc_tril' in var and 'Qc' not in var:
        var['Qc'] = _gram_view(var['sQc_tril'])
    return var


class _StaleVariables:
    """Problem mixin: ``variables()`` also answers to the stale names."""

    #: variables the older code stored transposed (orthonormal COLUMNS:
    #: hfb320_sqrt_zoh.py:123 assigns ``np.eye(2*nx, nx)`` to ``pred_orth``;
    #: HEAD has orthonormal rows, symfem.py:145-149); handed out as
    #: transposed, writable views of the same storage
    stale_transposed = ()

    def variables(self, dvec):
        var = stale_variable_aliases(super().variables(dvec))
        for name in self.stale_transposed:
            if name in var:
                var[name] = np.swapaxes(var[name], -1, -2)
        return var


def _stale_classes():
    """Classes the stale scripts name that HEAD no longer has, under their
    original names (the generated-class name derives from it)."""
    return {
        # hfb320_sqrt_zoh.py:97,101 -> DiscretizedNoise + ZOH
        # (attas_sp_ml_ndisc.py:49,53)
        'NaturalSqrtZOHModel': type(
            'NaturalSqrtZOHModel',
            (models.DiscretizedNoiseModel, models.ZOHDynamicsModel), {}),
        'NaturalSqrtZOHProblem': type(
            'NaturalSqrtZOHProblem',
            (_StaleVariables, problems.DiscretizedNoiseProblem,
             problems.ZOHDynamicsProblem),
            {'stale_transposed': ('pred_orth', 'corr_orth')}),
        # blackbox_innov_bal.py:32,55,59 -> BalancedDT
        'InnovationBalDTModel': type('InnovationBalDTModel',
                                     (models.BalancedDTModel,), {}),
        'InnovationBalDTProblem': type(
            'InnovationBalDTProblem',
            (_StaleVariables, problems.BalancedDTProblem), {}),
    }


def add_stale_aliases():
    for name, cls in _stale_classes().items():
        setattr(models if name.endswith('Model') else problems, name, cls)


class GeneratedModuleFinder(importlib.abc.MetaPathFinder,
                            importlib.abc.Loader):
    """Resolves ``import <Class>_nx<a>_nu<b>_ny<c>`` -- the module a script's
    ``get_model`` expects from an earlier ``save_generated_model`` run
    (hfb320_sqrt_zoh.py:40-58, blackbox_innov_bal.py:21-39) -- by generating
    it on the fly from ``symfem.<Class>``.  Installed LAST on
    ``sys.meta_path``: a module file the script wrote itself wins."""

    PATTERN = re.compile(r'^(?P<cls>[A-Za-z_]\w*Model)_nx(?P<nx>\d+)'
                         r'_nu(?P<nu>\d+)_ny(?P<ny>\d+)$')

    def find_spec(self, name, path=None, target=None):
        m = self.PATTERN.match(name)
        symfem = sys.modules.get('symfem')
        if not m or symfem is None or not hasattr(symfem, m['cls']):
            return None
        return importlib.machinery.ModuleSpec(name, self)

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        m = self.PATTERN.match(module.__name__)
        cls = getattr(sys.modules['symfem'], m['cls'])
        symmodel = cls(nx=int(m['nx']), nu=int(m['nu']), ny=int(m['ny']))
        exec(symmodel.print_code(), module.__dict__)


def install(mirrors=None):
    """Register the ``ceacoest`` / ``sym2num`` stand-ins (idempotent).

    ``mirrors``: bind ``symfem`` / ``fem`` to this package's mirrors (default:
    only if no module of that name is importable).
    """
    ceacoest = _module('ceacoest', optim=optim)
    modelling = _module('ceacoest.modelling', symoptim=symoptim)
    ceacoest.modelling = modelling
    ceacoest.__path__ = []
    modelling.__path__ = []
    sym2num = _module('sym2num')
    sym2num.__path__ = []
    sym2num.model = _module('sym2num.model')
    for name, mod in (('ceacoest', ceacoest), ('ceacoest.optim', optim),
                      ('ceacoest.modelling', modelling),
                      ('ceacoest.modelling.symoptim', symoptim),
                      ('sym2num', sym2num), ('sym2num.model', sym2num.model)):
        existing = sys.modules.get(name)
        if existing is None or getattr(existing, '__cfem_compat__', False) \
                or existing in (optim, symoptim):
            sys.modules[name] = mod
    add_stale_aliases()
    if not any(isinstance(f, GeneratedModuleFinder) for f in sys.meta_path):
        sys.meta_path.append(GeneratedModuleFinder())
    if mirrors is None:
        import importlib.util
        mirrors = importlib.util.find_spec('symfem') is None
    if mirrors:
        sys.modules['symfem'] = models
        sys.modules['fem'] = problems
