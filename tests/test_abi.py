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


def test_merge_peaks_host_utility():
    """gr_merge_peaks (host code, no device involved): per-rank lists in chromosome order, every chromosome in exactly
    one of them -> one list in chromosome order; runs of every length, empty lists, a single list."""
    import numpy as np
    from genrich_b200.dist import PEAK_DTYPE
    api = capi.load_cuda()
    rng = np.random.RandomState(4)
    for nlists, nchrom in ((1, 5), (3, 40), (8, 25), (4, 3)):
        owner = rng.randint(0, nlists, nchrom)
        per = [[] for _ in range(nlists)]
        want = []
        for c in range(nchrom):
            n = int(rng.choice([0, 1, 2, 3, 7, 64, 1000]))
            rec = np.zeros(n, PEAK_DTYPE)
            rec["chrom"] = c
            rec["start"] = np.sort(rng.randint(0, 1 << 30, n))
            rec["end"] = rec["start"] + 100
            per[owner[c]].append(rec)
            want.append(rec)
        lists = [np.concatenate(p) if p else np.zeros(0, PEAK_DTYPE) for p in per]
        want = np.concatenate(want)
        out = np.zeros(len(want) + 1, PEAK_DTYPE)
        ptrs = (C.c_void_p * nlists)(*[l.ctypes.data for l in lists])
        cnts = (C.c_uint64 * nlists)(*[len(l) for l in lists])
        assert api.merge_peaks(ptrs, cnts, nlists, out.ctypes.data_as(C.c_void_p)) == 0
        assert out[:len(want)].tobytes() == want.tobytes()
