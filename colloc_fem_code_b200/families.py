"""The class compositions the reference scripts build, by short name.

Each entry lists the model bases and problem bases in the order the scripts
write them (cooperative multiple inheritance, as in the reference):

===============  =====================================================
``innovation``   attas_sp_innov.py:84-86
``balanced``     attas_sp_innov_bal.py:85-87
``ml``           attas_sp_ml.py:85-87
``ml_zoh``       attas_sp_ml_zoh.py:49-54
``ndisc_zoh``    attas_sp_ml_ndisc.py:49-54
``ml_balanced``  mc_blackbox_cfem.py:25-30
``trapezoid``    extension: trapezoidal collocation, no reference class
===============  =====================================================
"""

from . import models, problems

COMPOSITIONS = {
    'innovation': (('InnovationDTModel',), ('InnovationDTProblem',)),
    'balanced': (('BalancedDTModel',), ('BalancedDTProblem',)),
    'ml': (('MaximumLikelihoodDTModel',), ('MaximumLikelihoodDTProblem',)),
    'ml_zoh': (('MaximumLikelihoodDTModel', 'ZOHDynamicsModel'),
               ('MaximumLikelihoodDTProblem', 'ZOHDynamicsProblem')),
    'ndisc_zoh': (('DiscretizedNoiseModel', 'ZOHDynamicsModel'),
                  ('DiscretizedNoiseProblem', 'ZOHDynamicsProblem')),
    'ml_balanced': (('MaximumLikelihoodDTModel', 'BalancedDTModel'),
                    ('MaximumLikelihoodDTProblem', 'BalancedDTProblem')),
    # extension without reference counterpart (trapezoidal collocation)
    'trapezoid': (('TrapezoidalCTModel',), ('TrapezoidalCTProblem',)),
}

_model_classes = {}


def _compose(module, names, label):
    bases = tuple(getattr(module, n) for n in names)
    if len(bases) == 1:
        return bases[0]
    return type(label, bases, {})


def model_class(kind, nx, nu, ny):
    """Compiled model class of a family at given dimensions (cached: the
    symbolic differentiation is the slow part, as in the reference)."""
    key = (kind, nx, nu, ny)
    if key not in _model_classes:
        sym = _compose(models, COMPOSITIONS[kind][0], 'Model')
        sym = type(f'{kind}_nx{nx}_nu{nu}_ny{ny}', (sym,), {})
        _model_classes[key] = sym(nx=nx, nu=nu, ny=ny).compile_class()
    return _model_classes[key]


def problem_class(kind):
    return _compose(problems, COMPOSITIONS[kind][1], 'Problem')


def make_problem(kind, y, u, nx, dt=None):
    nu, ny = u.shape[1], y.shape[1]
    model = model_class(kind, nx, nu, ny)()
    if dt is not None:
        model.dt = dt
    return problem_class(kind)(model, y, u)
