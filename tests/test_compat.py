"""The reference's own modules run unchanged on top of this package
(INTEGRATION.md section 1a).  Host logic only -- no GPU."""

import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(ref)r)              # running from the reference's directory
import colloc_fem_code_b200.compat as compat
compat.install()
import symfem, fem                      # the reference's files, unchanged
assert symfem.__file__.startswith(%(ref)r)
class Model(symfem.MaximumLikelihoodDTModel, symfem.ZOHDynamicsModel):
    generated_name = 'GeneratedMLZOH'
class Problem(fem.MaximumLikelihoodDTProblem, fem.ZOHDynamicsProblem):
    pass
g = np.load(%(golden)r)
nx, nu, ny = (int(v) for v in g['dims'])
model = Model(nx=nx, nu=nu, ny=ny).compile_class()()
assert type(model).__name__ == 'GeneratedMLZOH'
model.dt = float(g['dt'])
p = Problem(model, g['y'], g['u'])
assert (p.ndec, p.ncons) == (int(g['ndec']), int(g['ncons']))
jr, jc = p.constr_jac_ind(); hr, hc = p.lag_hess_ind()
assert (jr == g['jac_row']).all() and (jc == g['jac_col']).all()
assert (hr == g['hess_row']).all() and (hc == g['hess_col']).all()
# print_code round trip (mc_blackbox_cfem.py:81-95)
code = Model(nx=nx, nu=nu, ny=ny).print_code()
ns = {}
exec(code, ns)
assert ns['GeneratedMLZOH']().nx == nx
print('ok')
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'symfem.py')),
                    reason='reference tree only exists in the build container')
def test_reference_modules_run_on_the_package():
    golden = os.path.join(ROOT, 'tests', 'golden', 'ml_zoh_nx2_nu1_ny2_N5.npz')
    code = SCRIPT % {'root': ROOT, 'ref': REF, 'golden': golden}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.strip().endswith('ok')


def test_mirrors_and_stale_aliases():
    code = r'''
import sys
sys.path.insert(0, %r)
import colloc_fem_code_b200.compat as compat
compat.install(mirrors=True)
import symfem, fem
import numpy as np
m = symfem.InnovationBalDTModel(nx=2, nu=1, ny=2).compile_class()()
p = fem.InnovationBalDTProblem(m, np.zeros((5, 2)), np.zeros((5, 1)))
assert 'sW_diag' in p.decision
assert symfem.tril_mat(2, np.arange(3.0)).shape == (2, 2)      # stale 2-arg form
assert symfem.tril_mat(np.arange(3.0))[1, 0] == 1.0
assert hasattr(fem, 'NaturalSqrtZOHProblem')
print('ok')
''' % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
