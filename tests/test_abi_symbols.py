"""CPU: the C-ABI library loads and exports every symbol include/mv2d_b200.h declares (no
compute calls), and the ctypes mirrors match the header's struct sizes."""
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'mv2d_b200.h')).read()
    return sorted(set(re.findall(r'MV2D_API\s+[\w\s\*]+?\b(mv2d_\w+)\s*\(', text)))


def test_header_symbols_exported_and_bound():
    from mv2d_b200 import lib
    names = _declared()
    assert len(names) >= 15
    handle = lib.load()   # verifies ABI version and struct sizes, raises otherwise
    bound = {n for n, _, _ in lib.SYMBOLS}
    for n in names:
        assert hasattr(handle, n), f'{n} not exported by libmv2d_b200.so'
        assert n in bound, f'{n} has no ctypes prototype in mv2d_b200/lib.py'
    assert bound <= set(names), f'bound but undeclared: {bound - set(names)}'


def test_argument_errors_do_not_launch():
    """Error convention: <0 and a message, nothing launched (no GPU needed for these)."""
    from mv2d_b200 import lib
    h = lib.load()
    before = h.mv2d_launch_count()
    assert h.mv2d_pe3d(None, None) < 0
    assert b'pe3d' in h.mv2d_last_error()
    assert h.mv2d_gemm(None, 0, None, 0, None, None, 0, 1, 1, 1, 0, None) < 0
    assert h.mv2d_geom_prep(None, 6, None, None, None) < 0
    assert h.mv2d_launch_count() == before


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under mv2d_b200/ may import it."""
    pkg = os.path.join(ROOT, 'mv2d_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f
