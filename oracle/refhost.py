"""Import the reference's own ``symfem.py`` / ``fem.py`` UNCHANGED on top of the
oracle engine (only possible where /root/reference exists, i.e. in the build
container).  TEST INFRASTRUCTURE ONLY."""

import importlib
import os
import sys

REFERENCE_DIR = os.environ.get('CFEM_REFERENCE_DIR', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'symfem.py'))


def load():
    """Return the reference's (symfem, fem) modules bound to the oracle."""
    if not available():
        raise RuntimeError(f'no reference tree at {REFERENCE_DIR}')
    here = os.path.dirname(os.path.abspath(__file__))
    shim = os.path.join(here, 'shim')
    root = os.path.dirname(here)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in list(sys.modules)
                  if k.split('.')[0] in ('ceacoest', 'sym2num', 'fem',
                                         'symfem')}
    for k in saved_mods:
        del sys.modules[k]
    sys.path[:0] = [shim, root, REFERENCE_DIR]
    try:
        symfem = importlib.import_module('symfem')
        fem = importlib.import_module('fem')
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k.split('.')[0] in ('ceacoest', 'sym2num', 'fem', 'symfem'):
                del sys.modules[k]
        sys.modules.update({k: m for k, m in saved_mods.items() if m})
    return symfem, fem


#: reference class compositions per problem kind (same keys as
#: oracle.ref_models.KINDS); the tuples are (model bases, problem bases) in the
#: order the reference scripts list them.
_COMPOSITIONS = {
    'innovation': (('InnovationDTModel',), ('InnovationDTProblem',)),
    'balanced': (('BalancedDTModel',), ('BalancedDTProblem',)),
    'ml': (('MaximumLikelihoodDTModel',), ('MaximumLikelihoodDTProblem',)),
    'ml_zoh': (('MaximumLikelihoodDTModel', 'ZOHDynamicsModel'),
               ('MaximumLikelihoodDTProblem', 'ZOHDynamicsProblem')),
    'ndisc_zoh': (('DiscretizedNoiseModel', 'ZOHDynamicsModel'),
                  ('DiscretizedNoiseProblem', 'ZOHDynamicsProblem')),
    'ml_balanced': (('MaximumLikelihoodDTModel', 'BalancedDTModel'),
                    ('MaximumLikelihoodDTProblem', 'BalancedDTProblem')),
}

_loaded = None
_models = {}


def make_problem(kind, y, u, nx, dt=None):
    """Problem of the reference's own classes (fem.py / symfem.py unchanged)."""
    global _loaded
    if _loaded is None:
        _loaded = load()
    symfem, fem = _loaded
    nu, ny = u.shape[1], y.shape[1]
    mnames, pnames = _COMPOSITIONS[kind]
    key = (kind, nx, nu, ny)
    if key not in _models:
        mbases = tuple(getattr(symfem, n) for n in mnames)
        mcls = mbases[0] if len(mbases) == 1 else type('Model', mbases, {})
        _models[key] = mcls(nx=nx, nu=nu, ny=ny).compile_class()
    model = _models[key]()
    if dt is not None:
        model.dt = dt
    pbases = tuple(getattr(fem, n) for n in pnames)
    pcls = pbases[0] if len(pbases) == 1 else type('Problem', pbases, {})
    return pcls(model, y, u)
