"""CPU tests: the C-ABI libraries load and export every symbol include/heifcuda.h declares; the
CUDA-free build refuses to reconstruct (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import heif_b200 as hb
from heif_b200 import _lib
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "heifcuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(n for n, _, _ in _lib.SYMBOLS)


@pytest.mark.parametrize("host_only", [True, False])
def test_library_exports_every_declared_symbol(host_only):
    path = _lib.lib_path(host_only)
    if not os.path.exists(path):
        pytest.skip(path + " not built")
    L = C.CDLL(path)
    for name in declared_symbols():
        assert hasattr(L, name), name


def test_host_only_build_has_no_fallback():
    L = _lib.load(True)
    assert L.hc_has_cuda_engine() == 0
    assert not L.hc_engine_create(0)
    assert b"no CPU fallback" in L.hc_last_error()


def test_csc_select_mirrors_reference_table():
    # 8-bit 4:2:0 full range -> fixed point; limited range / other chroma -> fp32
    p = hb.csc_select(6, 1, 1, 1, 8, 0, hb.OUT_RGB, host_only=True)
    assert (p.mode, p.r_cr_i, p.g_cb_i, p.g_cr_i, p.b_cb_i) == (0, 359, -88, -183, 454)
    assert hb.csc_select(2, 2, 0, 1, 8, 0, hb.OUT_RGB, host_only=True).mode == 1
    assert hb.csc_select(1, 1, 1, 2, 8, 0, hb.OUT_RGB, host_only=True).mode == 1
    assert hb.csc_select(1, 1, 1, 1, 10, 0, hb.OUT_RRGGBB_LE, host_only=True).mode == 1
    assert hb.csc_select(0, 1, 1, 3, 8, 0, hb.OUT_RGB, host_only=True).mode == 2
    with pytest.raises(hb.HeifCudaError):
        hb.csc_select(11, 1, 1, 1, 8, 0, hb.OUT_RGB, host_only=True)
