"""The generated C-ABI library builds, loads and exports every symbol that
include/cfem.h declares (no compute calls: there is no GPU here)."""

import ctypes
import json
import os
import re

import numpy as np

from colloc_fem_code_b200 import backend, families

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'cfem.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(cfem_[a-z0-9_]+)\s*\(', text))


def _library(kind='innovation', dims=(2, 1, 2)):
    nx, nu, ny = dims
    p = families.make_problem(kind, np.zeros((4, ny)), np.zeros((4, nu)), nx,
                              dt=0.1)
    return p, backend.Library.for_structure(p.structure)


def test_header_and_binding_agree():
    assert _header_symbols() == set(backend.ABI)


def test_library_exports_every_symbol():
    _, lib = _library()
    dll = ctypes.CDLL(lib.path)
    for name in _header_symbols():
        assert hasattr(dll, name), name
    assert lib.cfem_abi_version() == 1


def test_model_sizes_match_host_layout(golden):
    nx, nu, ny = golden['dims']
    p = families.make_problem(golden['kind'], golden['y'], golden['u'], nx,
                              dt=golden['dt'])
    if (golden['kind'], golden['dims']) not in (
            ('innovation', (2, 1, 2)), ('ml', (2, 1, 2)),
            ('ndisc_zoh', (2, 1, 2))):
        return      # keep the CPU suite short: three libraries are enough
    lib = backend.Library.for_structure(p.structure)
    out = [ctypes.c_int64() for _ in range(4)]
    rc = lib.cfem_model_sizes(golden['N'], 0, *[ctypes.byref(o) for o in out])
    assert rc == 0
    assert [o.value for o in out] == [p.ndec, p.ncons, p.nnzjac, p.nnzhess]
    model = json.loads(lib.cfem_model_json().decode())
    assert [v['name'] for v in model['vars']] == list(p.decision)
    assert lib.cfem_model_sizes(1, 0, None, None, None, None) != 0


def test_create_fails_loudly_without_a_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    p, lib = _library()
    with pytest.raises(backend.CfemError):
        p.obj(np.zeros(p.ndec))


def test_kernel_options_generate_and_build():
    """The two alternative forms of the per-sample kernel (fence-free
    retirement of the partial sums, tile loads before the prologue;
    profiles/r02_stall_breakdown.md) stay buildable and export the same ABI.
    Their results are checked on the GPU by running the parity suite with
    CFEM_RETIRE=flag CFEM_EARLY_LOADS=1."""
    from colloc_fem_code_b200 import codegen
    p = families.make_problem('innovation', np.zeros((4, 1)), np.zeros((4, 1)),
                              1, dt=0.1)
    base = codegen.generate(p.structure)['main']
    assert 'tree_reduce<' in base and 'publish_partial<' not in base
    assert 'kFlagRetire = false' in base
    src = codegen.generate(p.structure, retire='flag', early_loads=1)['main']
    kernels = src[src.index('// mask 1:'):]
    assert 'publish_partial<' in kernels and 'tree_reduce<' not in kernels
    assert 'collect_partials(' in src and 'kFlagRetire = true' in src
    # the first tile's loads come before the parameter staging
    k31 = kernels[kernels.index('cfem_sample_kernel_m31'):]
    assert k31.index('stage_rows_async') < k31.index('stage_contig')
    path = backend.build_library(p.structure,
                                 backend.structure_label(p.structure),
                                 retire='flag', early_loads=1)
    dll = ctypes.CDLL(path)
    for name in _header_symbols():
        assert hasattr(dll, name), name
