"""ctypes binding of the C-ABI declared in include/genrich_cuda.h.

The binding is generic over (shared library, symbol prefix) so that the parity
tests can drive the CPU oracle (prefix ``orc_``, built from oracle/) through the
very same Python surface as the CUDA library (prefix ``gr_``).  Nothing in this
package ever loads the oracle: :func:`load_cuda` is the only loader the product
uses, and it raises if ``libgenrich_cuda.so`` is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(HERE, "csrc", "libgenrich_cuda.so")

GR_SKIP = -1.0


class GrChrom(C.Structure):
    _fields_ = [("len", C.c_uint32), ("skip", C.c_uint8), ("owned", C.c_uint8),
                ("reserved", C.c_uint16)]


class GrParams(C.Structure):
    _fields_ = [("min_pqval", C.c_float), ("qval_opt", C.c_int32),
                ("min_auc", C.c_float), ("min_len", C.c_int32),
                ("max_gap", C.c_int32), ("keep_pileups", C.c_int32),
                ("genome_len", C.c_uint64)]


class GrSampleStats(C.Structure):
    _fields_ = [("frag_len", C.c_double), ("ctrl_frag", C.c_double),
                ("lambda_", C.c_float), ("factor", C.c_float),
                ("genome_len", C.c_uint64), ("n_expt", C.c_uint64),
                ("n_ctrl", C.c_uint64), ("n_pval", C.c_uint64),
                ("n_clamped", C.c_uint64)]


class GrPeak(C.Structure):
    _fields_ = [("chrom", C.c_int32), ("summit", C.c_uint32),
                ("start", C.c_int64), ("end", C.c_int64), ("auc", C.c_float),
                ("pval", C.c_float), ("qval", C.c_float), ("reserved", C.c_float)]


class GrRunStats(C.Structure):
    _fields_ = [("genome_len", C.c_uint64), ("n_peaks", C.c_uint64),
                ("peak_bp", C.c_uint64), ("n_intervals", C.c_uint64),
                ("n_distinct_p", C.c_uint64), ("all_q_one", C.c_int32),
                ("n_replicates", C.c_int32)]


class GrStageTime(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_double),
                ("launches", C.c_uint64), ("bytes", C.c_uint64)]


PEAK_DTYPE = np.dtype([("chrom", "<i4"), ("summit", "<u4"), ("start", "<i8"),
                       ("end", "<i8"), ("auc", "<f4"), ("pval", "<f4"),
                       ("qval", "<f4"), ("reserved", "<f4")])

# every symbol include/genrich_cuda.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "gr_create", "gr_destroy", "gr_set_exclusions", "gr_excluded_bp", "gr_set_params", "gr_reset", "gr_strerror",
    "gr_last_error_detail", "gr_sample_begin", "gr_push_intervals",
    "gr_push_intervals_device", "gr_prefetch_intervals", "gr_push_packed", "gr_prefetch_packed",
    "gr_pack6_layout", "gr_push_packed6", "gr_prefetch_packed6",
    "gr_sample_pileup", "gr_sample_skipped", "gr_sample_sums", "gr_replicate_finish", "gr_replicate_finish_device",
    "gr_sums_device", "gr_stream", "gr_replicate_stats",
    "gr_replicate_end", "gr_pvalues_finalize", "gr_bh_local_hist",
    "gr_bh_set_global", "gr_bh_local_hist_host", "gr_bh_set_global_host", "gr_load_pvalues", "gr_call_peaks", "gr_call_peaks_enqueue", "gr_call_peaks_done", "gr_peaks_device", "gr_merge_peaks", "gr_fetch_intervals",
    "gr_timing_enable", "gr_timing_get", "gr_timing_reset",
    "gr_kernel_launches", "gr_scan_form", "gr_synchronize", "gr_timer_start", "gr_timer_stop",
    "gr_pinned_alloc", "gr_pinned_free",
]

STATUS_TEXT = {
    0: "ok", 1: "bad argument or call order", 2: "CUDA failure",
    3: "Cannot allocate memory", 4: ": read aligned beyond reference end",
    5: "Experimental sample has no analyzable fragments",
    6: "No analyzable genome (length=0)", 7: "Invalid pileup value (< 0)",
    8: "Disallowed number of alignments", 9: "interval on unknown/unowned chromosome",
    10: "Invalid df in pchisq()", 11: "Genome length does not match p-value length",
    12: "no CUDA device", 13: "more than 32767 fragment starts/ends on one base (int16 saturation, Genrich.c:2558)",
}


class GenrichError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = STATUS_TEXT.get(status, "unknown")
        super().__init__(f"{where}: status {status} ({msg}) {detail}".strip())


@dataclass
class Intervals:
    """One RLE array of the reference's ``Pileup`` type (Genrich.h:173-176)."""
    end: np.ndarray
    val: np.ndarray
    expt: np.ndarray | None = None
    ctrl: np.ndarray | None = None


class Api:
    """Thin functional wrapper of one shared library exporting <prefix>* symbols."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found - build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.has_device = prefix == "gr_"
        L, p = self.lib, prefix
        vp, i32, u64, dbl = C.c_void_p, C.c_int32, C.c_uint64, C.c_double

        def fn(name, res, args):
            f = getattr(L, p + name)
            f.restype = res
            f.argtypes = args
            return f

        if self.has_device:
            self.create = fn("create", C.c_int, [C.POINTER(vp), C.POINTER(GrChrom), i32, C.POINTER(GrParams), i32])
            self.reset = fn("reset", C.c_int, [vp])
            self.strerror = fn("strerror", C.c_char_p, [C.c_int])
            self.last_error_detail = fn("last_error_detail", C.c_char_p, [vp])
            self.push_intervals_device = fn("push_intervals_device", C.c_int, [vp, vp, u64])
            self.prefetch_intervals = fn("prefetch_intervals", C.c_int, [vp, vp, u64])
            self.sample_sums = fn("sample_sums", C.c_int, [vp, C.POINTER(dbl), C.POINTER(dbl)])
            self.replicate_finish_device = fn("replicate_finish_device", C.c_int, [vp, i32, u64])
            self.sums_device = fn("sums_device", C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)])
            self.stream = fn("stream", vp, [vp])
            self.replicate_stats = fn("replicate_stats", C.c_int, [vp, i32, C.POINTER(GrSampleStats)])
            self.push_packed = fn("push_packed", C.c_int, [vp, vp, u64])
            self.prefetch_packed = fn("prefetch_packed", C.c_int, [vp, vp, u64])
            self.pack6_layout = fn("pack6_layout", C.c_int, [vp, vp])
            self.push_packed6 = fn("push_packed6", C.c_int, [vp, vp, u64])
            self.prefetch_packed6 = fn("prefetch_packed6", C.c_int, [vp, vp, u64])
            self.peaks_device = fn("peaks_device", C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)])
            self.call_peaks_enqueue = fn("call_peaks_enqueue", C.c_int, [vp, C.POINTER(vp), C.POINTER(u64)])
            self.call_peaks_done = fn("call_peaks_done", C.c_int, [vp, vp, C.POINTER(i32), C.POINTER(GrRunStats)])
            self.bh_local_hist_host = fn("bh_local_hist_host", C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)])
            self.bh_set_global_host = fn("bh_set_global_host", C.c_int, [vp, vp, vp, u64, u64])
            self.merge_peaks = fn("merge_peaks", C.c_int, [C.POINTER(vp), C.POINTER(u64), i32, vp])
            self.timing_enable = fn("timing_enable", C.c_int, [vp, i32])
            self.timing_get = fn("timing_get", C.c_int, [vp, C.POINTER(GrStageTime), i32, C.POINTER(i32)])
            self.timing_reset = fn("timing_reset", C.c_int, [vp])
            self.kernel_launches = fn("kernel_launches", u64, [vp])
            self.scan_form = fn("scan_form", C.c_int, [vp, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64)])
            self.synchronize = fn("synchronize", C.c_int, [vp])
            self.timer_start = fn("timer_start", C.c_int, [vp])
            self.timer_stop = fn("timer_stop", C.c_int, [vp, C.POINTER(dbl)])
        else:
            self.create = fn("create", C.c_int, [C.POINTER(vp), C.POINTER(GrChrom), i32, C.POINTER(GrParams)])
        self.destroy = fn("destroy", None, [vp])
        self.set_params = fn("set_params", C.c_int, [vp, C.POINTER(GrParams)])
        self.load_pvalues = fn("load_pvalues", C.c_int, [vp, vp, vp, vp, vp, u64])
        self.set_exclusions = fn("set_exclusions", C.c_int, [vp, vp, vp, vp, u64])
        self.excluded_bp = fn("excluded_bp", C.c_int, [vp, vp])
        self.sample_begin = fn("sample_begin", C.c_int, [vp, i32, vp])
        self.push_intervals = fn("push_intervals", C.c_int, [vp, vp, u64])
        self.sample_pileup = fn("sample_pileup", C.c_int, [vp, C.POINTER(dbl)])
        self.sample_skipped = fn("sample_skipped", C.c_int, [vp, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)])
        self.replicate_finish = fn("replicate_finish", C.c_int, [vp, dbl, dbl, i32, u64, C.POINTER(GrSampleStats)])
        self.replicate_end = fn("replicate_end", C.c_int, [vp, C.POINTER(GrSampleStats)])
        self.pvalues_finalize = fn("pvalues_finalize", C.c_int, [vp])
        self.bh_local_hist = fn("bh_local_hist", C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)])
        self.bh_set_global = fn("bh_set_global", C.c_int, [vp, vp, vp, u64, u64])
        self.call_peaks = fn("call_peaks", C.c_int, [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(GrRunStats)])
        self.fetch_intervals = fn("fetch_intervals", C.c_int,
                                  [vp, i32, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)])


_libm = C.CDLL("libm.so.6")
_libm_log10f = _libm.log10f
_libm_log10f.restype = C.c_float
_libm_log10f.argtypes = [C.c_float]

_cuda_api: Api | None = None


def load_cuda() -> Api:
    """Load libgenrich_cuda.so.  There is no fallback: a missing library is fatal."""
    global _cuda_api
    if _cuda_api is None:
        _cuda_api = Api(CUDA_LIB, "gr_")
    return _cuda_api


def _as_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _np_from(ptr, n, dtype):
    if not ptr or n == 0:
        return np.empty(0, dtype=dtype)
    # one memmove into a fresh array: np.frombuffer(...).copy() of a structured dtype copies field by
    # field (1.3 ms for 51 k peak records where the memmove takes 0.09 ms)
    out = np.empty(n, dtype=dtype)
    C.memmove(out.ctypes.data, ptr, n * np.dtype(dtype).itemsize)
    return out


class Context:
    """One engine context (= the reference's Chrom[] plus peak parameters)."""

    def __init__(self, api: Api, chrom_len, params: GrParams, device: int = 0,
                 skip=None, owned=None):
        self.api = api
        n = len(chrom_len)
        arr = (GrChrom * n)()
        for i in range(n):
            arr[i].len = int(chrom_len[i])
            arr[i].skip = int(skip[i]) if skip is not None else 0
            arr[i].owned = int(owned[i]) if owned is not None else 1
        self.nchrom = n
        self.chrom_len = np.asarray(chrom_len, dtype=np.uint32)
        self._h = C.c_void_p()
        if api.has_device:
            rc = api.create(C.byref(self._h), arr, n, C.byref(params), device)
        else:
            rc = api.create(C.byref(self._h), arr, n, C.byref(params))
        self._check(rc, "create")

    def _check(self, rc, where):
        if rc != 0:
            detail = ""
            if self.api.has_device and self._h:
                d = self.api.last_error_detail(self._h)
                detail = d.decode() if d else ""
            raise GenrichError(rc, where, detail)

    def close(self):
        if self._h:
            self.api.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self.api.reset(self._h), "reset")

    def set_exclusions(self, regions):
        """regions: iterable of (chromosome index, start, end) -- the -E BED records, unmerged.
        Call before the first sample."""
        r = np.asarray(list(regions), dtype=np.int64).reshape(-1, 3)
        c = np.ascontiguousarray(r[:, 0], dtype=np.int32)
        a = np.ascontiguousarray(r[:, 1], dtype=np.uint32)
        b = np.ascontiguousarray(r[:, 2], dtype=np.uint32)
        self._check(self.api.set_exclusions(self._h, _as_ptr(c), _as_ptr(a), _as_ptr(b), len(c)), "set_exclusions")

    def excluded_bp(self) -> np.ndarray:
        """Excluded bp per chromosome after merging / clamping (saveXBed 1144)."""
        out = np.zeros(self.nchrom, dtype=np.uint64)
        self._check(self.api.excluded_bp(self._h, _as_ptr(out)), "excluded_bp")
        return out

    # -- seam IN ---------------------------------------------------------------
    def sample_begin(self, is_ctrl: bool, save=None):
        sv = None
        if save is not None:
            sv = np.ascontiguousarray(save, dtype=np.uint8)
        self._check(self.api.sample_begin(self._h, int(is_ctrl), _as_ptr(sv) if sv is not None else None),
                    "sample_begin")

    def push_intervals(self, recs: np.ndarray):
        recs = np.ascontiguousarray(recs, dtype=np.int32).reshape(-1, 4)
        self._check(self.api.push_intervals(self._h, _as_ptr(recs), recs.shape[0]), "push_intervals")

    def push_intervals_device(self, dptr: int, n: int):
        self._check(self.api.push_intervals_device(self._h, C.c_void_p(dptr), n), "push_intervals_device")

    def prefetch_ptr(self, host_ptr: int, n: int):
        self._check(self.api.prefetch_intervals(self._h, C.c_void_p(host_ptr), n), "prefetch_intervals")

    def push_packed(self, recs: np.ndarray):
        """recs: uint64 GR_PACK records (host.pack_records)."""
        recs = np.ascontiguousarray(recs, dtype=np.uint64)
        self._check(self.api.push_packed(self._h, _as_ptr(recs), recs.shape[0]), "push_packed")

    def push_packed_ptr(self, ptr: int, n: int):
        """ptr: host (pinned or not) or device address of n GR_PACK records."""
        self._check(self.api.push_packed(self._h, C.c_void_p(ptr), n), "push_packed")

    def prefetch_packed_ptr(self, host_ptr: int, n: int):
        self._check(self.api.prefetch_packed(self._h, C.c_void_p(host_ptr), n), "prefetch_packed")

    def pack6_layout(self):
        """Cell offset of every chromosome in this context's layout (uint64; all-ones = not held here),
        or None when the layout does not fit the 6-byte record format."""
        off = np.zeros(self.nchrom, dtype=np.uint64)
        if self.api.pack6_layout(self._h, _as_ptr(off)) != 0:
            return None
        return off

    def push_packed6(self, recs: np.ndarray):
        recs = np.ascontiguousarray(recs, dtype=np.uint16).reshape(-1, 3)
        self._check(self.api.push_packed6(self._h, _as_ptr(recs), recs.shape[0]), "push_packed6")

    def push_packed6_ptr(self, ptr: int, n: int):
        self._check(self.api.push_packed6(self._h, C.c_void_p(ptr), n), "push_packed6")

    def prefetch_packed6_ptr(self, host_ptr: int, n: int):
        self._check(self.api.prefetch_packed6(self._h, C.c_void_p(host_ptr), n), "prefetch_packed6")

    def push_ptr(self, host_ptr: int, n: int):
        self._check(self.api.push_intervals(self._h, C.c_void_p(host_ptr), n), "push_intervals")

    def sample_pileup(self) -> np.ndarray:
        sums = np.zeros(self.nchrom, dtype=np.float64)
        self._check(self.api.sample_pileup(self._h, sums.ctypes.data_as(C.POINTER(C.c_double))), "sample_pileup")
        return sums

    def sample_skipped(self, is_ctrl=False):
        """saveInterval 2558-2573: (dropped for overflow, dropped for underflow, array of (arrival index << 1 | underflow))
        of the last experimental / control sample."""
        a, b, n, p = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_void_p()
        self._check(self.api.sample_skipped(self._h, int(is_ctrl), C.byref(a), C.byref(b), C.byref(p), C.byref(n)), "sample_skipped")
        return a.value, b.value, _np_from(p.value, n.value, np.uint64)

    def replicate_finish(self, frag_len, ctrl_frag, has_ctrl, genome_len=0) -> GrSampleStats:
        st = GrSampleStats()
        self._check(self.api.replicate_finish(self._h, frag_len, ctrl_frag, int(has_ctrl),
                                              int(genome_len), C.byref(st)), "replicate_finish")
        return st

    # -- the same without host round trips (CUDA library only) ---------------------
    def sample_pileup_async(self):
        self._check(self.api.sample_pileup(self._h, None), "sample_pileup")

    def sample_sums(self, want_ctrl=True):
        e = np.zeros(self.nchrom, dtype=np.float64)
        c = np.zeros(self.nchrom, dtype=np.float64) if want_ctrl else None
        dp = C.POINTER(C.c_double)
        self._check(self.api.sample_sums(self._h, e.ctypes.data_as(dp), c.ctypes.data_as(dp) if want_ctrl else None),
                    "sample_sums")
        return e, c

    def replicate_finish_device(self, has_ctrl, genome_len=0):
        self._check(self.api.replicate_finish_device(self._h, int(has_ctrl), int(genome_len)), "replicate_finish_device")

    def sums_device_ptrs(self):
        e, c = C.c_void_p(), C.c_void_p()
        self._check(self.api.sums_device(self._h, C.byref(e), C.byref(c)), "sums_device")
        return e.value, c.value

    def stream_handle(self) -> int:
        return int(self.api.stream(self._h) or 0)

    def replicate_stats(self, replicate: int) -> GrSampleStats:
        st = GrSampleStats()
        self._check(self.api.replicate_stats(self._h, replicate, C.byref(st)), "replicate_stats")
        return st

    def replicate_end(self) -> GrSampleStats:
        st = GrSampleStats()
        self._check(self.api.replicate_end(self._h, C.byref(st)), "replicate_end")
        return st

    # -- peaks -----------------------------------------------------------------
    def set_params(self, params: GrParams):
        self._check(self.api.set_params(self._h, C.byref(params)), "set_params")

    def load_pvalues(self, chrom_start, end, pval, qval=None):
        """-P (callPeaksLog 1277): the final -log10 p (and q) intervals come from the caller."""
        cs = np.ascontiguousarray(chrom_start, dtype=np.uint64)
        e = np.ascontiguousarray(end, dtype=np.uint32)
        p = np.ascontiguousarray(pval, dtype=np.float32)
        q = np.ascontiguousarray(qval, dtype=np.float32) if qval is not None else None
        self._check(self.api.load_pvalues(self._h, _as_ptr(cs), _as_ptr(e), _as_ptr(p), _as_ptr(q) if q is not None else None,
                                          len(e)), "load_pvalues")

    def pvalues_finalize(self):
        self._check(self.api.pvalues_finalize(self._h), "pvalues_finalize")

    def bh_local_hist_ptrs(self):
        k, l, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.api.bh_local_hist(self._h, C.byref(k), C.byref(l), C.byref(n)), "bh_local_hist")
        return k.value, l.value, n.value

    def bh_set_global_ptrs(self, keys_ptr, lens_ptr, n, genome_len):
        self._check(self.api.bh_set_global(self._h, C.c_void_p(keys_ptr), C.c_void_p(lens_ptr), n, genome_len),
                    "bh_set_global")

    def call_peaks(self, to_host=True):
        """to_host=False (CUDA library): the records stay on the device (peaks_device_ptr)."""
        p, n, st = C.c_void_p(), C.c_uint64(), GrRunStats()
        self._check(self.api.call_peaks(self._h, C.byref(p) if to_host else None, C.byref(n), C.byref(st)), "call_peaks")
        if not to_host:
            return np.empty(0, PEAK_DTYPE), st
        return _np_from(p.value, n.value, PEAK_DTYPE), st

    def call_peaks_enqueue(self):
        """First half of gr_call_peaks for launchers that gather: (device address of the slot, record capacity)."""
        p, cap = C.c_void_p(), C.c_uint64()
        self._check(self.api.call_peaks_enqueue(self._h, C.byref(p), C.byref(cap)), "call_peaks_enqueue")
        return p.value, cap.value

    def call_peaks_done(self, header_ptr: int):
        """Second half: this context's 64-byte slot header as the host sees it -> (redo, run stats)."""
        redo, st = C.c_int32(), GrRunStats()
        self._check(self.api.call_peaks_done(self._h, C.c_void_p(header_ptr), C.byref(redo), C.byref(st)), "call_peaks_done")
        return redo.value, st

    def peaks_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.api.peaks_device(self._h, C.byref(p), C.byref(n)), "peaks_device")
        return p.value or 0, n.value

    def fetch(self, which: int, replicate: int, chrom: int) -> Intervals | None:
        e, v, x, c, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.api.fetch_intervals(self._h, which, replicate, chrom, C.byref(e), C.byref(v),
                                             C.byref(x), C.byref(c), C.byref(n)), "fetch_intervals")
        if not e.value:
            return None
        return Intervals(_np_from(e.value, n.value, np.uint32), _np_from(v.value, n.value, np.float32),
                         _np_from(x.value, n.value, np.float32) if x.value else None,
                         _np_from(c.value, n.value, np.float32) if c.value else None)

    # -- device-only helpers -----------------------------------------------------
    def timing(self, on=True):
        self._check(self.api.timing_enable(self._h, int(on)), "timing_enable")

    def timing_get(self):
        arr = (GrStageTime * 64)()
        n = C.c_int32()
        self._check(self.api.timing_get(self._h, arr, 64, C.byref(n)), "timing_get")
        return {arr[i].name.decode(): (arr[i].ms, arr[i].launches, arr[i].bytes) for i in range(n.value)}

    def timing_reset(self):
        self._check(self.api.timing_reset(self._h), "timing_reset")

    def kernel_launches(self) -> int:
        return int(self.api.kernel_launches(self._h))

    def scan_form(self):
        """(form, hot_entries, entries) of the last sample: 0 rank form, 1 CTA form, -1 not bucketed"""
        f, h, e = C.c_int32(), C.c_uint64(), C.c_uint64()
        self._check(self.api.scan_form(self._h, C.byref(f), C.byref(h), C.byref(e)), "scan_form")
        return f.value, h.value, e.value

    def synchronize(self):
        self._check(self.api.synchronize(self._h), "synchronize")

    def timer_start(self):
        self._check(self.api.timer_start(self._h), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.api.timer_stop(self._h, C.byref(ms)), "timer_stop")
        return ms.value


def make_params(p=None, q=None, min_auc=200.0, min_len=0, max_gap=100,
                keep_pileups=False, genome_len=0) -> GrParams:
    """Thresholds as getArgs() converts them (Genrich.c:5815-5817): -log10f in float."""
    qopt = q is not None
    thr = C.c_float(q if qopt else (0.01 if p is None else p))
    pq = -_libm_log10f(thr)      # glibc log10f on a float, like the reference
    return GrParams(float(pq), int(qopt), float(min_auc), int(min_len), int(max_gap),
                    int(keep_pileups), int(genome_len))
