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
are aliased on the mirrors to their nearest HEAD equivalents.
"""

import sys
import types

from . import models, optim, problems, symoptim


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__dict__['__cfem_compat__'] = True
    return mod


class _NaturalSqrtZOHModel(models.DiscretizedNoiseModel,
                           models.ZOHDynamicsModel):
    """hfb320_sqrt_zoh.py:97 -> DiscretizedNoise + ZOH (attas_sp_ml_ndisc.py:49)."""


class _NaturalSqrtZOHProblem(problems.DiscretizedNoiseProblem,
                             problems.ZOHDynamicsProblem):
    """hfb320_sqrt_zoh.py:101."""


def add_stale_aliases():
    models.NaturalSqrtZOHModel = _NaturalSqrtZOHModel
    problems.NaturalSqrtZOHProblem = _NaturalSqrtZOHProblem
    # blackbox_innov_bal.py:32,55,59
    models.InnovationBalDTModel = models.BalancedDTModel
    problems.InnovationBalDTProblem = problems.BalancedDTProblem


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
    if mirrors is None:
        import importlib.util
        mirrors = importlib.util.find_spec('symfem') is None
    if mirrors:
        sys.modules['symfem'] = models
        sys.modules['fem'] = problems
