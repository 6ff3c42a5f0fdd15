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


def test_function_level_indices_compose_to_the_problem_level(golden):
    """The per-function operator interface (adfem.py:20-120: jac_ind /
    hess_ind with variable-local indices) plus the problem's offsets gives
    exactly the global COO index arrays."""
    p = _problem(golden)
    specs = {**p.decision, **p.dependent}
    shapes = {n: s.shape for n, s in specs.items()}
    rows, cols = [], []
    for name, reg in p.constraints.items():
        fun = getattr(p.model, name)
        ind = fun.jac_ind(shapes, reg.block.shape)
        assert fun.jac_nnz(shapes, reg.block.shape) == \
            sum(v.shape[1] for v in ind.values())
        for (wrt,), (w, o) in ind.items():
            rows.append(reg.block.offset + o)
            cols.append(specs[wrt].offset + w)
    jr, jc = p.constr_jac_ind()
    np.testing.assert_array_equal(np.concatenate(rows), jr)
    np.testing.assert_array_equal(np.concatenate(cols), jc)
    hrow, hcol = [], []
    regs = list(p.objectives.items()) + list(p.constraints.items())
    for name, reg in regs:
        fun = getattr(p.model, name)
        for (w0, w1), (i0, i1, o) in fun.hess_ind(shapes,
                                                  reg.block.shape).items():
            a, b = specs[w0].offset + i0, specs[w1].offset + i1
            hrow.append(np.maximum(a, b))
            hcol.append(np.minimum(a, b))
    hr, hc = p.lag_hess_ind()
    np.testing.assert_array_equal(np.concatenate(hrow), hr)
    np.testing.assert_array_equal(np.concatenate(hcol), hc)
    import pytest
    with pytest.raises(RuntimeError):
        p.model.dynamics.jac_val()
