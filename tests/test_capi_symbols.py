"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares; the ctypes
table mirrors the header; argument validation that needs no GPU works."""
import ctypes
import os
import re

from util import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lagomorph_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lgm_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "lagomorph_b200", "liblagomorph_b200.so"))
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_ctypes_table_matches_header(lm):
    from lagomorph_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib.lgm_version() == 1


def test_no_torch_in_abi():
    out = os.popen("ldd %s" % os.path.join(ROOT, "lagomorph_b200", "liblagomorph_b200.so")).read()
    assert "torch" not in out and "python" not in out.lower()


def test_argument_validation_without_gpu(lm):
    from lagomorph_b200 import _lib as L
    sh = L.shape_arr((4, 4))
    # bad dim
    rc = L.lib.lgm_interp_fwd(0, None, None, None, 1, 1, 1, 5, sh, 1.0, None)
    assert rc == -1 and b"two- and three-dimensional" in L.lib.lgm_last_error()
    # bad dtype
    rc = L.lib.lgm_jtvf_fwd(7, None, None, None, 1, 2, 2, sh, 0, 0, None)
    assert rc == -1
    # thin dimension
    rc = L.lib.lgm_jtvf_fwd(0, None, None, None, 1, 2, 2, L.shape_arr((4, 1)), 0, 0, None)
    assert rc == -1 and b"thin" in L.lib.lgm_last_error()
    # workspace sizes: fast path (one spectrum buffer) and direct-DFT path (two)
    assert L.lib.lgm_fluid_workspace_bytes(0, 2, 3, L.shape_arr((16, 16, 16))) == 2 * 3 * 16 * 16 * 9 * 8
    assert L.lib.lgm_fluid_workspace_bytes(1, 1, 2, L.shape_arr((3, 3))) == 2 * (1 * 2 * 3 * 2 * 16)
    assert L.lib.lgm_epdiff_scratch_bytes(0, 1, 3, L.shape_arr((16, 16, 16))) >= 3 * 16 ** 3 * 4
