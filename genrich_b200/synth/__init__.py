"""Python face of the deterministic workload generator (gen_synth.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgrsynth.so")
BIN = os.path.join(os.path.dirname(HERE), "bin", "gen_synth")


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("nchrom", C.c_int32), ("chrom_len", C.POINTER(C.c_uint32)),
                ("nfrag", C.c_uint64), ("enrich", C.c_double), ("peak_spacing", C.c_uint32),
                ("peak_sigma", C.c_double), ("frag_min", C.c_int32), ("frag_max", C.c_int32),
                ("read_len", C.c_int32), ("multimap_frac", C.c_double), ("multimap_max", C.c_int32)]


def build(force: bool = False) -> None:
    src = os.path.join(HERE, "gen_synth.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if force or not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-DGEN_SYNTH_MAIN", "-o", BIN, src, "-lm"])


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.synth_count_records.restype = C.c_uint64
        _lib.synth_count_records.argtypes = [C.POINTER(SynthParams)]
        _lib.synth_fragments.restype = C.c_uint64
        _lib.synth_fragments.argtypes = [C.POINTER(SynthParams), C.c_uint64, C.c_uint64, C.c_void_p]
    return _lib


class Workload:
    """A seeded synthetic sample: chromosome table + templates."""

    def __init__(self, chrom_len, nfrag, seed, enrich=0.2, spacing=50000, sigma=150.0,
                 fmin=100, fmax=400, multimap=0.0, mmax=12, read_len=50):
        self.chrom_len = np.ascontiguousarray(chrom_len, dtype=np.uint32)
        self.p = SynthParams(seed, len(self.chrom_len),
                             self.chrom_len.ctypes.data_as(C.POINTER(C.c_uint32)), nfrag, enrich,
                             spacing, sigma, fmin, fmax, read_len, multimap, mmax)

    def fragments(self, first: int = 0, n: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        lib = _load()
        if n is None:
            n = self.p.nfrag - first
        cap = n * (1 if self.p.multimap_frac <= 0 else 10)
        if out is None:
            out = np.empty((cap, 4), dtype=np.int32)
        w = lib.synth_fragments(C.byref(self.p), first, n, out.ctypes.data_as(C.c_void_p))
        return out[:w]

    def cli_args(self) -> list[str]:
        a = ["--lens", ",".join(str(int(x)) for x in self.chrom_len), "--frags", str(self.p.nfrag),
             "--seed", str(self.p.seed), "--enrich", repr(self.p.enrich), "--spacing", str(self.p.peak_spacing),
             "--sigma", repr(self.p.peak_sigma), "--fmin", str(self.p.frag_min), "--fmax", str(self.p.frag_max)]
        if self.p.multimap_frac > 0:
            a += ["--multimap", repr(self.p.multimap_frac), "--mmax", str(self.p.multimap_max)]
        return a

    def write_sam(self, path: str) -> None:
        build()
        subprocess.check_call([BIN] + self.cli_args() + ["--sam", path])
