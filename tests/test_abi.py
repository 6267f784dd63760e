"""CPU: the C-ABI shared library builds, loads and exports every symbol include/refnerf_b200.h declares
(no compute calls without a GPU), and the product has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from refnerf_pl_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'refnerf_b200.h')).read()
    declared = set(re.findall(r'RN_API\s+[\w\s\*]+?\b(rn_\w+)\s*\(', hdr))
    assert len(declared) >= 20
    from refnerf_pl_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_metadata_calls(lib):
    from refnerf_pl_b200 import _lib
    from oracle import refnerf_oracle as O
    assert lib.rn_abi_version() == 3
    names = _lib.param_names()
    assert len(names) == 46 and names[0] == 'spatial_net.0.weight' and names[-1] == 'rgb.bias'
    shapes = O.param_shapes()
    total = 0
    for i, nme in enumerate(names):
        base, kind = nme.rsplit('.', 1)
        n, k = shapes[base]
        want = n * k if kind == 'weight' else n
        assert lib.rn_mlp_param_numel(i) == want, nme
        total += want
    assert total == 1110158            # parameter count of the Ref-NeRF NerfMLP
    for prec in (0, 1, 2, 3):
        assert lib.rn_mlp_packed_bytes(prec) > 1_000_000
    cfg = _lib.RnMlpConfig(1, 1, 1, 0.5, -1.0, 1.0, 0.0, 0.001, 4096, 0)
    assert lib.rn_mlp_workspace_bytes(ctypes.byref(cfg), 1) > lib.rn_mlp_workspace_bytes(ctypes.byref(cfg), 0) > 0


def test_no_cpu_fallback():
    from refnerf_pl_b200 import ops
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.resample(torch.rand(2, 3), torch.rand(2, 2), torch.zeros(2, 1), torch.ones(2, 1), 8, 0.01, 1.0, 0.0, 1.0, False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'refnerf_pl_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('the oracle', '').replace('oracle/', ''), f
