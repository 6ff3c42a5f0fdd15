"""The table form of the parameter-only constraints (codegen.py): the parser of
the printed expressions and the tables it builds, emulated on the CPU.

The device loop multiplies coefficient and factors left to right and adds the
terms in the printed order with __dmul_rn / __dadd_rn, i.e. plain IEEE double
operations -- exactly what Python floats do.  Evaluating the tables here must
therefore give the SAME BITS as evaluating the printed expression itself."""

import numpy as np
import pytest

from colloc_fem_code_b200 import codegen, families, symoptim

G = codegen.Generator


def test_parser_forms():
    assert G._parse_sum('1.0') == [(1.0, [], None)]
    assert G._parse_sum('-v_a_0') == [(-1.0, ['v_a_0'], None)]
    assert G._parse_sum('v_a_0 - 0.5*v_b_1*v_b_1 + 2e-3*v_c_0') == [
        (1.0, ['v_a_0'], None), (-0.5, ['v_b_1', 'v_b_1'], None),
        (0.002, ['v_c_0'], None)]
    t = G._parse_sum('v_a_0 - (1.0/3.0)*v_b_1*(0.5*v_a_0 + v_d_2) - 1.0*v_e_0')
    assert t[1] == (-(1.0 / 3.0), ['v_b_1'],
                    [(0.5, ['v_a_0'], None), (1.0, ['v_d_2'], None)])
    # anything else keeps generated code
    for code in ('v_a_0*(v_b_0 + v_c_0)*v_d_0',         # group not last
                 'v_a_0*(v_b_0*(v_c_0 + v_e_0))',       # deeper nesting
                 'v_a_0/v_b_0', 'sqrt(v_a_0)', 'v_a_0*2.0', '(v_a_0 + v_b_0)'):
        assert G._parse_sum(code) is None, code
    assert G._parse_flat('v_a_0*(v_b_0 + v_c_0)') is None


def _emulate(tab, e, value_of, lam_of, sigma):
    def product(t):
        p = tab['coeff'][t]
        for k in range(tab['fac0'][t], tab['fac0'][t + 1]):
            p = p * value_of(tab['fac'][k])
        return p
    total = 0.0
    for t in range(tab['term0'][e], tab['term0'][e + 1]):
        p = product(t)
        if tab['in0'][t] >= 0:
            inner = 0.0
            for u in range(tab['in0'][t], tab['in1'][t]):
                inner = inner + product(u)
            p = p * inner
        total = total + p
    m = tab['mult'][e]
    if m == 0xFFFFFFFE:
        total = sigma * total
    elif m != 0xFFFFFFFF:
        total = lam_of(m >> 20, m & 0xFFFFF) * total
    return total


@pytest.mark.parametrize('kind,dims', [('ml', (2, 1, 2)),
                                       ('ml_zoh', (2, 1, 2)),
                                       ('ml_balanced', (5, 3, 3)),
                                       ('ndisc_zoh', (4, 2, 3))])
def test_tables_reproduce_the_printed_expressions_bit_for_bit(kind, dims):
    nx, nu, ny = dims
    p = families.make_problem(kind, np.zeros((4, ny)), np.zeros((4, nu)), nx,
                              dt=0.1)
    gen = G(p.structure)
    gen.sources()
    tab = gen.param_table
    assert gen.n_param_table + gen.n_param_code == gen.n_param_entries
    assert gen.n_param_table >= 0.9 * gen.n_param_entries
    rng = np.random.default_rng(0)
    st = p.structure
    var_vals = [rng.normal(size=max(1, v['core'] * 4)) for v in st.vars]
    scal_vals = rng.uniform(0.05, 0.2, size=max(1, len(st.scalars)))
    lam_vals = rng.normal(size=(len(st.funs), 4096))
    sigma = 0.7

    def value_of(f):
        space, idx, flat = f >> 30, (f >> 20) & 1023, f & 0xFFFFF
        return float(var_vals[idx][flat] if space == 0 else scal_vals[idx])

    def lam_of(ci, i):
        return float(lam_vals[ci][i])

    checked = 0
    for e, ent in enumerate(tab['entries']):
        fun = ent['fun']
        env = {}
        for a_, fl in ent['deps']:
            ref = fun['args'][a_]
            env[symoptim.c_ident(a_, fl)] = float(
                var_vals[ref[1]][fl] if ref[0] == 'param'
                else scal_vals[ref[1]])
        direct = eval(ent['code'], {'__builtins__': {}}, env)
        if ent['mult'] is not None:
            mult = sigma if ent['mult'][0] == 'sigma' else \
                lam_of(ent['mult'][1], ent['mult'][2])
            direct = mult * direct
        got = _emulate(tab, e, value_of, lam_of, sigma)
        assert got == direct or (np.isnan(got) and np.isnan(direct)), \
            (e, ent['code'][:80], got, direct)
        # destination packing
        d = tab['dest'][e]
        assert (d >> 30, (d >> 20) & 1023, d & 0xFFFFF) == \
            (ent['kind'], ent['blk'], ent['off'])
        checked += 1
    assert checked == gen.n_param_table


def test_parser_on_random_expressions():
    """Random sums of products in the printed style (signs, literal
    coefficients, repeated symbols, one trailing group): evaluating the parsed
    form left to right equals Python's evaluation of the string, bit for bit."""
    rng = np.random.default_rng(5)
    names = [f'v_p{k}_{j}' for k in range(3) for j in range(4)]
    env = {n: float(rng.normal()) for n in names}

    def rand_term(group_ok):
        parts = []
        if rng.random() < 0.6:
            parts.append(rng.choice(['0.5', '2.0', '1.0', '0.33333333333333331',
                                     '(1.0/3.0)', '1e-3', '12.0']))
        for _ in range(int(rng.integers(0 if parts else 1, 5))):
            parts.append(str(rng.choice(names)))
        if not parts:
            parts.append(str(rng.choice(names)))
        if group_ok and rng.random() < 0.4 and \
                not parts[-1].startswith(('0', '1', '2', '(')):
            inner = rand_sum(False)
            parts.append('(' + inner + ')')
        return '*'.join(parts)

    def rand_sum(group_ok):
        n = int(rng.integers(1, 6))
        out = ('-' if rng.random() < 0.3 else '') + rand_term(group_ok)
        for _ in range(n - 1):
            out += (' + ' if rng.random() < 0.5 else ' - ') + rand_term(group_ok)
        return out

    def evaluate(terms):
        total = 0.0
        for c, ids, inner in terms:
            p = c
            for i in ids:
                p = p * env[i]
            if inner is not None:
                p = p * evaluate(inner)
            total = total + p
        return total

    parsed = 0
    for _ in range(400):
        code = rand_sum(True)
        terms = G._parse_sum(code)
        if terms is None:
            continue
        parsed += 1
        assert evaluate(terms) == eval(code, {'__builtins__': {}}, env), code
    assert parsed > 300
