"""Host logic of the product package against the golden fixtures (no GPU):
decision / constraint layout (fem.py:36-57 and subclasses) and the COO index
arrays, which must equal the oracle's entry for entry, in order."""

import numpy as np

from colloc_fem_code_b200 import families


def _problem(g):
    nx, nu, ny = g['dims']
    return families.make_problem(g['kind'], g['y'], g['u'], nx, dt=g['dt'])


def test_layout_matches_reference(golden):
    p = _problem(golden)
    assert (p.ndec, p.ncons) == (golden['ndec'], golden['ncons'])
    lay = golden['layout']
    assert [[n, list(s.shape), s.offset] for n, s in p.decision.items()] \
        == lay['decision']
    assert [[n, list(s.shape), s.offset] for n, s in p.dependent.items()] \
        == lay['dependent']
    assert [[n, list(r.block.shape), r.block.offset]
            for n, r in p.constraints.items()] == lay['constraints']
    assert [[n, list(r.block.shape)] for n, r in p.objectives.items()] \
        == lay['objectives']


def test_index_arrays_match_oracle(golden):
    p = _problem(golden)
    jr, jc = p.constr_jac_ind()
    hr, hc = p.lag_hess_ind()
    assert p.nnzjac == len(golden['jac_val']) == len(jr)
    assert p.nnzhess == len(golden['hess_val']) == len(hr)
    np.testing.assert_array_equal(jr, golden['jac_row'])
    np.testing.assert_array_equal(jc, golden['jac_col'])
    np.testing.assert_array_equal(hr, golden['hess_row'])
    np.testing.assert_array_equal(hc, golden['hess_col'])
    assert (hr >= hc).all()


def test_views_are_writable(golden):
    p = _problem(golden)
    dvec = np.zeros(p.ndec)
    var = p.variables(dvec)
    var['x'][:] = 1.0
    var['A'][0, 0] = 7.0
    assert dvec[p.decision['A'].offset] == 7.0
    assert dvec.sum() == 7.0 + p.decision['x'].size
    np.testing.assert_array_equal(var['xnext'], var['x'][1:])
    cvec = np.zeros(p.ncons)
    p.unpack_constraints(cvec)['innovation'][:] = 2.0
    assert cvec.sum() == 2.0 * p.constraints['innovation'].block.size
