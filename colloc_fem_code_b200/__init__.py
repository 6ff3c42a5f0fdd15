"""B200-native evaluation of the collocation-based filter-error method.

A from-scratch implementation of the data-parallel hot path of
dimasad/colloc-fem-code: the per-sample evaluation behind the IPOPT callback
set (objective, gradient, constraints, sparse Jacobian values, Lagrangian
Hessian values) of the problems in the reference's ``fem.py`` / ``symfem.py``.

Layout of the package (only what the path needs):

``symoptim``  symbolic model front-end (boundary of ``ceacoest.modelling.symoptim``)
``models``    model families, mirror of the reference's ``symfem.py``
``optim``     problem glue + sparsity indices (boundary of ``ceacoest.optim``)
``problems``  problem families, mirror of the reference's ``fem.py``
``codegen``   CUDA code generator (replaces sym2num's NumPy generation)
``backend``   nvcc build + ctypes binding of the C ABI in ``include/cfem.h``
``csrc``      hand-written device helpers and host side of the C ABI
"""

from . import symoptim, optim, models, problems  # noqa: F401

__all__ = ['symoptim', 'optim', 'models', 'problems']
