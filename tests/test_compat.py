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


STALE_RUNNER = r'''
import os, sys, warnings
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import colloc_fem_code_b200.compat as compat
compat.install(mirrors=True)        # the stale classes live on the mirrors
from colloc_fem_code_b200 import nlp, synthetic
from oracle import ref_models
from nlp_helpers import OracleEvaluator

# No GPU in this test: the callbacks come from the CPU oracle of the same
# problem family (the layouts are identical, tests/test_host_layout.py).
KIND = {'NaturalSqrtZOHProblem': 'ndisc_zoh', 'InnovationBalDTProblem': 'balanced'}
def cpu_evaluator(problem):
    o = ref_models.make_problem(KIND[type(problem).__name__], problem.y,
                                problem.u, problem.model.nx,
                                dt=getattr(problem.model, 'dt', 0.1))
    assert (o.ndec, o.ncons) == (problem.ndec, problem.ncons)
    return OracleEvaluator(o)
nlp.GpuEvaluator = cpu_evaluator
cap = nlp.Solver.add_int_option
def capped(self, key, value):       # the point is to execute, not to converge
    cap(self, key, min(value, 12) if key == 'max_iter' else value)
nlp.Solver.add_int_option = capped

src = open(%(script)r).read().split('\n')
main_at = next(i for i, l in enumerate(src) if l.startswith("if __name__ == '__main__':"))
ns = {'__name__': 'stale_script', '__file__': %(script)r}
exec(compile('\n'.join(src[:main_at]), %(script)r, 'exec'), ns)

nx, nu, ny, N = %(dims)r
exp = synthetic.experiment(3, N, nx, nu, ny)
files = {}
if %(which)r == 'hfb320':
    ns['load_data'] = lambda: (0.1 * np.arange(N), exp['u'], exp['y'])
else:
    from colloc_fem_code_b200 import fit
    guess = fit.predictor_guess(exp['y'], exp['u'], exp['A'], exp['B'],
                                exp['C'], exp['D'], np.zeros((nx, ny)))
    sRp = np.zeros((ny, ny)); sRp[np.tril_indices(ny)] = guess['sRp_tril']
    files = {'u.txt': exp['u'], 'y.txt': exp['y'], 'a.txt': exp['A'],
             'b.txt': exp['B'], 'c.txt': exp['C'], 'd.txt': exp['D'],
             'k.txt': guess['Ln'], 'xpred.txt': guess['x'],
             'epred.txt': guess['en'], 'gram.txt': np.full(nx, 0.5),
             'isRp.txt': np.linalg.inv(sRp)}
    class NP:                       # np with the data files patched in
        def __getattr__(self, name):
            return getattr(np, name)
        def loadtxt(self, path, *a, **k):
            return np.array(files[os.path.basename(path)])
    ns['np'] = NP()
    ns['load_data'] = lambda: (exp['u'], exp['y'])

first, last = %(lines)r            # 1-based, inclusive: the cited statement range
body = src[main_at + 1:last]
body = ['if True:'] + body
with warnings.catch_warnings(record=True) as caught:
    warnings.simplefilter('always')
    exec(compile('\n'.join(body), %(script)r + ':main', 'exec'), ns)
assert ns['problem'].ndec == len(ns['decopt'])
assert np.all(np.isfinite(ns['decopt']))
assert ns['eopt'].shape == (N, ny) and ns['L'].shape == (nx, ny)
print('warnings:', sorted({str(w.message)[:60] for w in caught}))
print('status:', ns['info']['status'], 'iterations:', ns['info'].get('iterations'))
%(extra)s
print('ok')
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'symfem.py')),
                    reason='reference tree only exists in the build container')
@pytest.mark.parametrize('which,script,dims,lines,extra', [
    ('hfb320', 'hfb320_sqrt_zoh.py', (4, 2, 7, 40), (97, 204),
     "assert ns['Qc'].shape == (4, 4) and ns['isRp'].shape == (7, 7)\n"
     "assert type(ns['model']).__name__ == 'GeneratedNaturalSqrtZOHModel'"),
    ('blackbox', 'blackbox_innov_bal.py', (5, 3, 3, 60), (55, 118),
     "assert ns['W'].shape == (5, 5) and np.all(np.diag(ns['W']) >= 0)\n"
     "assert np.all(np.isfinite(ns['isRp']))"),
])
def test_stale_script_bodies_execute(which, script, dims, lines, extra):
    """hfb320_sqrt_zoh.py:97-204 and blackbox_innov_bal.py:55-118 -- class names
    and variables of an older parametrisation -- run as written through
    ``compat`` (stale classes, variable aliases, on-demand generated-model
    modules), with the data loaders patched to synthetic arrays and the CPU
    oracle serving the callbacks."""
    code = STALE_RUNNER % {'root': ROOT, 'script': os.path.join(REF, script),
                           'dims': dims, 'which': which, 'lines': lines,
                           'extra': extra}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, timeout=900, cwd='/tmp')
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert out.stdout.strip().endswith('ok')


def test_stale_variable_views():
    """L / e alias Ln / en; W_diag, isRp_tril, Qc are derived views."""
    from colloc_fem_code_b200 import compat, models, problems
    compat.add_stale_aliases()
    m = models.NaturalSqrtZOHModel(nx=2, nu=1, ny=2).compile_class()()
    m.dt = 0.1
    p = problems.NaturalSqrtZOHProblem(m, np.zeros((6, 2)), np.zeros((6, 1)))
    dec = np.zeros(p.ndec)
    var = p.variables(dec)
    var['L'][:] = 2.0
    var['e'][:] = 3.0
    assert (var['Ln'] == 2.0).all() and (p.variables(dec)['en'] == 3.0).all()
    var['isRp_tril'][models.tril_diag(2)] = 1e2     # hfb320_sqrt_zoh.py:118
    np.testing.assert_allclose(p.variables(dec)['sRp_tril'], [1e-2, 0, 1e-2])
    var['Qc'][:] = np.eye(2) * 1e-2                 # hfb320_sqrt_zoh.py:125
    np.testing.assert_allclose(p.variables(dec)['sQc_tril'], [0.1, 0, 0.1])
    np.testing.assert_allclose(np.asarray(p.variables(dec)['Qc']),
                               np.eye(2) * 1e-2)
    # a bounds vector: nothing to transfer, the HEAD block keeps its value
    lo = np.full(p.ndec, -np.inf)
    with pytest.warns(UserWarning, match='older parametrisation'):
        p.variables(lo)['isRp_tril'][models.tril_diag(2)] = 0
    assert np.isinf(p.variables(lo)['sRp_tril']).all()
    mb = models.InnovationBalDTModel(nx=2, nu=1, ny=2).compile_class()()
    pb = problems.InnovationBalDTProblem(mb, np.zeros((6, 2)),
                                         np.zeros((6, 1)))
    decb = np.zeros(pb.ndec)
    pb.variables(decb)['W_diag'][:] = [4.0, 9.0]    # blackbox_innov_bal.py:72
    np.testing.assert_allclose(pb.variables(decb)['sW_diag'], [2.0, 3.0])
    lob = np.full(pb.ndec, -np.inf)
    pb.variables(lob)['W_diag'][:] = 0              # blackbox_innov_bal.py:80
    assert (pb.variables(lob)['sW_diag'] == 0).all()


HEAD_RUNNER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
sys.path.insert(0, %(ref)r)              # the script runs from the reference's directory
import colloc_fem_code_b200.compat as compat
compat.install()                        # the one line a maintainer adds
from colloc_fem_code_b200 import nlp
from oracle import ref_models
from nlp_helpers import OracleEvaluator, attas_like_experiment

# No GPU in this test: the callbacks come from the CPU oracle of the same
# problem family (identical layouts, tests/test_host_layout.py).
def cpu_evaluator(problem):
    o = ref_models.make_problem(%(kind)r, problem.y, problem.u, problem.model.nx,
                                dt=getattr(problem.model, 'dt', None))
    assert (o.ndec, o.ncons) == (problem.ndec, problem.ncons)
    return OracleEvaluator(o)
nlp.GpuEvaluator = cpu_evaluator
cap = nlp.Solver.add_int_option
def capped(self, key, value):       # the point is to execute, not to converge
    cap(self, key, min(value, 15) if key == 'max_iter' else value)
nlp.Solver.add_int_option = capped

N = 80
exp = attas_like_experiment(2, N)       # every state measured, as the scripts impose (C = I)
src = open(%(script)r).read()
main_at = src.index("if __name__ == '__main__':")
ns = {'__name__': 'head_script', '__file__': %(script)r}
exec(compile(src[:main_at], %(script)r, 'exec'), ns)
import fem, symfem
assert symfem.__file__.startswith(%(ref)r) and fem.__file__.startswith(%(ref)r)
ns['load_data'] = lambda: (0.05 * np.arange(N), exp['u'].copy(), exp['y'].copy(),
                           np.zeros(2), np.ones(2), np.zeros(1), np.ones(1))
body = 'if True:' + src[main_at + len("if __name__ == '__main__':"):]
exec(compile(body, %(script)r + ':main', 'exec'), ns)
p = ns['problem']
assert p.ndec == len(ns['decopt']) and np.all(np.isfinite(ns['decopt']))
assert ns['xopt'].shape == (N, 2) and ns['enopt'].shape == (N, 2)
assert ns['yopt'].shape == (N, 2) and ns['sRp'].shape == (2, 2)
assert ns['info']['iterations'] >= 1
%(extra)s
print('status:', ns['info']['status'], 'iterations:', ns['info']['iterations'])
print('ok')
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'symfem.py')),
                    reason='reference tree only exists in the build container')
@pytest.mark.parametrize('script,kind,extra', [
    ('attas_sp_innov.py', 'innovation', ''),
    ('attas_sp_innov_bal.py', 'balanced',
     "assert ns['W'].shape == (2, 2)"),
    ('attas_sp_ml.py', 'ml',
     "assert ns['Pp'].shape == (2, 2) and ns['Kn'].shape == (2, 2)"),
    ('attas_sp_ml_zoh.py', 'ml_zoh',
     "assert ns['Ac'].shape == (2, 2) and abs(ns['model'].dt - 0.05) < 1e-12"),
    ('attas_sp_ml_ndisc.py', 'ndisc_zoh',
     "assert ns['Qc'].shape == (2, 2) and ns['Bc'].shape == (2, 1)"),
])
def test_head_script_bodies_execute(script, kind, extra):
    """The five current ATTAS scripts of the reference run AS WRITTEN, from the
    first import to the last statement, on the reference's own unchanged
    ``symfem.py`` / ``fem.py`` with ``compat.install()`` in front (initial
    guesses, bounds, scaling, ``problem.ipopt(...)``, options, ``solve``,
    unpacking of the optimum); only ``load_data`` is patched to synthetic
    arrays (the flight-test files are not shipped) and the CPU oracle serves
    the callbacks (no GPU here)."""
    code = HEAD_RUNNER % {'root': ROOT, 'ref': REF, 'kind': kind,
                          'script': os.path.join(REF, script), 'extra': extra}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, timeout=900, cwd='/tmp')
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert out.stdout.strip().endswith('ok')


MC_RUNNER = r'''
import os, sys, tempfile
import numpy as np, scipy.io
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(ref)r)
import colloc_fem_code_b200.compat as compat
compat.install()
from colloc_fem_code_b200 import synthetic
work = tempfile.mkdtemp()
os.chdir(work)                          # get_model writes the generated module here
sys.path.insert(0, work)
edir = os.path.join(work, 'mc_experim')
os.makedirs(edir)
nx, nu, ny, N = 5, 3, 3, 500            # mc_data_gen.m:5-16
scipy.io.savemat(os.path.join(edir, 'config.mat'), {'nx': nx, 'nu': nu, 'ny': ny})
exp = synthetic.experiment(0, N, nx, nu, ny)
scipy.io.savemat(os.path.join(edir, 'exp001.mat'), {'u': exp['u'], 'y': exp['y']})
script = %(script)r
src = open(script).read()
main_at = src.index("if __name__ == '__main__':")
ns = {'__name__': 'mc_script', '__file__': script}
exec(compile(src[:main_at], script, 'exec'), ns)
sys.argv = ['mc_blackbox_cfem.py', edir]
body = 'if True:' + src[main_at + len("if __name__ == '__main__':"):]
try:
    exec(compile(body, script + ':main', 'exec'), ns)
except SystemExit:
    pass                                # mc_blackbox_cfem.py:120 stops after the first problem pair
else:
    raise AssertionError('the script is expected to stop at its raise SystemExit')
name = 'GeneratedBalancedMaximumLikelihoodModel_nx5_nu3_ny3'
assert os.path.isfile(os.path.join(work, name + '.py'))         # print_code round trip
assert type(ns['model']).__name__ == 'GeneratedBalancedMaximumLikelihoodModel'
assert ns['ye'].shape == (250, 3)                               # second half of the record
ml, inn = ns['ml_prob'], ns['in_prob']
assert (ml.ndec, ml.ncons) == (2353, 2285)                      # SURVEY appendix B
assert (inn.ndec, inn.ncons) == (88 + 250 * 8, 250 * 8 - 5)
assert set(inn.decision) < set(ml.decision)
print('ok')
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'symfem.py')),
                    reason='reference tree only exists in the build container')
def test_mc_script_executes():
    """mc_blackbox_cfem.py as written (argument parsing, config.mat, the
    generated-model module written by ``print_code`` and imported back, the
    split of the record, both problem objects) on an experiment directory
    made here after mc_data_gen.m; the script itself stops after constructing
    the first pair of problems (its ``estimate`` is a stub)."""
    code = MC_RUNNER % {'root': ROOT, 'ref': REF,
                        'script': os.path.join(REF, 'mc_blackbox_cfem.py')}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert out.stdout.strip().endswith('ok')
