"""The C-ABI library loads (no GPU needed) and exports every symbol the header declares."""
import ctypes as C
import os
import re

from genrich_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "genrich_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    lib = C.CDLL(capi.CUDA_LIB)
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_struct_sizes_match_header():
    assert C.sizeof(capi.GrChrom) == 8
    assert C.sizeof(capi.GrParams) == 32
    assert C.sizeof(capi.GrPeak) == 40
    assert C.sizeof(capi.GrSampleStats) == 64
    assert C.sizeof(capi.GrRunStats) == 48


def test_no_device_fails_loudly():
    """Without a GPU the product must refuse to run (there is no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        return
    api = capi.load_cuda()
    par = capi.make_params(p=0.01)
    try:
        capi.Context(api, [1000], par)
    except capi.GenrichError as e:
        assert e.status == 12
    else:
        raise AssertionError("context created without a device")
