"""The C-ABI shared library loads without a GPU and exports every symbol include/jacks_b200.h declares
(no compute calls here).  Also: compute entry points fail loudly, never fall back, when no device is visible."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def declared_symbols():
    hdr = open(os.path.join(REPO, 'include', 'jacks_b200.h')).read()
    return sorted(set(re.findall(r'JACKS_API\s+[\w\s\*]+?\b(jacks_\w+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    from jacks_b200 import _native
    assert os.path.exists(_native.LIB_PATH), 'run `python -m jacks_b200.build` (or __graft_entry__.build())'
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_defaults():
    from jacks_b200 import _native
    lib = _native.lib()
    assert lib.jacks_abi_version() == 2
    p = _native.default_params()
    assert (p.n_iter, p.apply_w_hp, p.tol, p.mu0_x, p.var0_x, p.mu0_w, p.var0_w, p.tau_prior_strength) == \
        (50, 0, 0.1, 1.0, 1.0, 0.0, 1e4, 0.5)                   # infer.py:18, :55
    with pytest.raises(TypeError):
        _native.default_params(no_such_parameter=1)


def test_no_device_means_error_not_fallback():
    from jacks_b200 import _native
    if _native.lib().jacks_device_count() > 0:
        pytest.skip('a CUDA device is visible')
    with pytest.raises(_native.JacksError, match='no CPU fallback'):
        _native.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'jacks_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.replace('no CPU fallback', ''), os.path.join(root, f)
