"""Pin the CPU oracle: (a) function by function against the reference's own code
(oracle/_ref/libref_funcs.so, present wherever `make -C oracle ref` ran), (b) end
to end against files the unmodified reference wrote (tests/golden)."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from cases import CASES


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_outputs(case):
    meta, gold_np = util.golden(case)
    ctx, res, par = util.run_case(util.oracle_api(), case)
    got = util.host.format_narrowpeak(res.peaks, util.names_of(case))
    assert got == gold_np                        # every column, byte for byte
    assert res.run_stats.n_peaks == meta["peaks"]
    assert res.run_stats.peak_bp == meta["peak_bp"]
    assert res.run_stats.genome_len == meta["genome_len"]
    assert bool(res.run_stats.all_q_one) == meta["all_q_one"]
    for st, lam in zip(res.sample_stats, meta["lambda"]):
        assert "%f" % st.lambda_ == "%f" % lam
    facs = [st.factor for st, (e, c) in zip(res.sample_stats, case.reps) if c is not None]
    for f, g in zip(facs, meta["factor"]):
        assert "%f" % f == "%f" % g
    log = util.log_lines(ctx, case, par)
    assert len(log) == meta["log_lines"]
    assert util.sha_lines(log) == meta["log_sha256"]     # the whole -f file
    pile = util.pile_lines(ctx, case)
    assert len(pile) == meta["pile_lines"]
    assert util.sha_lines(pile) == meta["pile_sha256"]   # the whole -k file
    assert sum(st.n_clamped for st in res.sample_stats) == meta["clamp_warnings"]


needs_ref = pytest.mark.skipif(not os.path.exists(util.REF_FUNCS),
                               reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def libs():
    ref = C.CDLL(util.REF_FUNCS)
    orc = util.oracle_api().lib
    ref.ref_calcPval.restype = C.c_float
    ref.ref_calcPval.argtypes = [C.c_float, C.c_float]
    orc.orc_calc_pval.restype = C.c_float
    orc.orc_calc_pval.argtypes = [C.c_float, C.c_float]
    ref.ref_pchisq.restype = C.c_double
    ref.ref_pchisq.argtypes = [C.c_double, C.c_int]
    orc.orc_pchisq.restype = C.c_double
    orc.orc_pchisq.argtypes = [C.c_double, C.c_int]
    ref.ref_multPval.restype = C.c_float
    ref.ref_multPval.argtypes = [C.c_void_p, C.c_int]
    orc.orc_mult_pval.restype = C.c_float
    orc.orc_mult_pval.argtypes = [C.c_void_p, C.c_int]
    orc.orc_units_to_val.restype = C.c_float
    orc.orc_units_to_val.argtypes = [C.c_int32]
    ref.ref_updateVal.restype = C.c_float
    ref.ref_updateVal.argtypes = [C.c_int16, C.c_uint8, C.POINTER(C.c_int32), C.POINTER(C.c_uint8)]
    ref.ref_diff_add.argtypes = [C.POINTER(C.c_int16), C.POINTER(C.c_uint8), C.c_int, C.c_int]
    return ref, orc


def _bits(x):
    return np.float32(x).view(np.uint32)


@needs_ref
def test_calc_pval_bitexact(libs):
    ref, orc = libs
    rng = np.random.default_rng(5)
    ex = np.concatenate([np.arange(0, 300) / 1.0, rng.integers(0, 120 * 400, 3000) / 120.0, [0.0, 1e-3, 1e4, 3e5]])
    ct = np.concatenate([[0.0, -1.0, 0.5, 1.0, 6.999, 7.0, 7.001, 25.02, 100.0, 1e4], rng.random(40) * 60])
    for c in ct:
        for e in ex[:: 7 if c > 1 else 1]:
            a, b = ref.ref_calcPval(e, c), orc.orc_calc_pval(e, c)
            assert _bits(a) == _bits(b), (e, c, a, b)


@needs_ref
def test_pchisq_and_fisher_bitexact(libs):
    ref, orc = libs
    rng = np.random.default_rng(6)
    for df in (4, 6, 8, 20, 100, 400):
        for x in np.concatenate([[1e-9, 0.3, 1.0, 1.9, 2.0, df - 2.0, df + 0.0, 5.0 * df, 3000.0], rng.random(200) * 4 * df]):
            a, b = ref.ref_pchisq(x, df), orc.orc_pchisq(x, df)
            assert a == b or (np.isnan(a) and np.isnan(b)), (x, df, a, b)
    for n in (2, 3, 5):
        for _ in range(500):
            v = (rng.random(n) * rng.choice([0.5, 5, 50])).astype(np.float32)
            v[rng.random(n) < 0.15] = -1.0
            v[rng.random(n) < 0.1] = 0.0
            a = ref.ref_multPval(v.ctypes.data, n)
            b = orc.orc_mult_pval(v.ctypes.data, n)
            assert _bits(a) == _bits(b), (v, a, b)


@needs_ref
def test_units_encoding_bitexact(libs):
    """Drive the reference's addFrac/subFrac/updateVal on random difference cells and
    check the integer-1/120 restatement reproduces the float value bit for bit, and
    that a cell is zero iff the integer is zero."""
    ref, orc = libs
    rng = np.random.default_rng(7)
    counts = [1, 2, 3, 4, 5, 6, 8, 10]
    for trial in range(300):
        L = 60
        cov = (C.c_int16 * (L + 1))()
        frac = (C.c_uint8 * (L + 1))()
        units = np.zeros(L + 1, dtype=np.int64)
        for _ in range(rng.integers(1, 80)):
            s = int(rng.integers(0, L))
            e = int(rng.integers(s, L + 1))
            k = int(rng.choice(counts))
            cp = C.cast(C.byref(cov, 2 * s), C.POINTER(C.c_int16))
            fp = C.cast(C.byref(frac, s), C.POINTER(C.c_uint8))
            ref.ref_diff_add(cp, fp, k, +1)
            cp = C.cast(C.byref(cov, 2 * e), C.POINTER(C.c_int16))
            fp = C.cast(C.byref(frac, e), C.POINTER(C.c_uint8))
            ref.ref_diff_add(cp, fp, k, -1)
            units[s] += 120 // k
            units[e] -= 120 // k
        rc, rf, run = C.c_int32(0), C.c_uint8(0), 0
        for j in range(L + 1):
            assert (cov[j] != 0 or frac[j] != 0) == (units[j] != 0)
            if cov[j] or frac[j]:
                v = ref.ref_updateVal(cov[j], frac[j], C.byref(rc), C.byref(rf))
                run += int(units[j])
                assert _bits(v) == _bits(orc.orc_units_to_val(run)), (trial, j, run)


@needs_ref
def test_qvalues_bitexact(libs):
    ref, _ = libs
    ref.ref_computeQval.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p]
    api = util.oracle_api()
    rng = np.random.default_rng(8)
    for trial in range(20):
        n = int(rng.integers(5, 4000))
        pool = (rng.random(int(rng.integers(2, 200))) * 12).astype(np.float32)
        p = rng.choice(pool, n).astype(np.float32)
        p[rng.random(n) < 0.3] = 0.0
        lens = rng.integers(1, 500, n).astype(np.uint32)
        end = np.cumsum(lens).astype(np.uint32)
        G = int(end[-1])
        q_ref = np.zeros(n, dtype=np.float32)
        ref.ref_computeQval(p.ctypes.data, end.ctypes.data, n, G, q_ref.ctypes.data)
        # the oracle's own BH (orc_bh_local_hist + orc_bh_set_global, run by orc_call_peaks) on the same intervals:
        # the p array is handed over as -P does (orc_load_pvalues), the q of EVERY interval read back (fetch 3)
        ctx = util.capi.Context(api, [G], util.capi.make_params(p=0.01))
        ctx.load_pvalues([0, n], end, p)
        ctx.set_params(util.capi.make_params(q=0.05))
        _, rs = ctx.call_peaks()
        q_orc = ctx.fetch(3, 0, 0).val
        assert rs.n_distinct_p == len(np.unique(p.view(np.uint32)))
        assert np.array_equal(q_orc.view(np.uint32), q_ref.view(np.uint32)), trial


@needs_ref
def test_cell_cov_matches_reference_cells(libs):
    """orc_cell_cov(N) == Diff.cov of the reference's own cell after any sequence of addFrac / subFrac /
    ++ / -- that sums to N 1/120ths -- also for negative cells (the end positions of intervals)."""
    ref, orc = libs
    orc.orc_cell_cov.restype = C.c_int32
    orc.orc_cell_cov.argtypes = [C.c_int64]
    rng = np.random.RandomState(11)
    counts = [1, 2, 3, 4, 5, 6, 8, 10]
    for trial in range(300):
        cov, frac, N = C.c_int16(0), C.c_uint8(0), 0
        bias = rng.uniform(0.2, 0.8)
        for step in range(400):
            cnt = counts[rng.randint(8)]
            sign = 1 if rng.uniform() < bias else -1
            ref.ref_diff_add(C.byref(cov), C.byref(frac), cnt, sign)
            N += sign * (120 // cnt)
            assert orc.orc_cell_cov(N) == cov.value, (trial, step, N, cov.value, frac.value)
    # around the limits the rule tests for
    for N in (32767 * 120 - 1, 32767 * 120, 32767 * 120 + 193, -32768 * 120, -32768 * 120 + 193, -32768 * 120 + 194):
        c = orc.orc_cell_cov(N)
        assert (c == 32767) == (32767 * 120 <= N) if N > 0 else True
    assert orc.orc_cell_cov(-32768 * 120 + 193) == -32768 and orc.orc_cell_cov(-32768 * 120 + 75) in (-32768, -32767)


def test_saturation_rule_matches_reference():
    """saveInterval 2558-2573 (intervals skipped once a delta counter sits at INT16_MAX / INT16_MIN, in arrival
    order): the oracle on tests/satcase.py == what the unmodified reference wrote for its SAM view -- narrowPeak
    byte for byte, the -f / -k text by hash, and the very alignments it reported as skipped."""
    import hashlib
    import json
    import satcase
    from genrich_b200 import capi, host
    meta = json.load(open(os.path.join(util.GOLDEN, "sat_hot.json")))
    recs = satcase.records()
    par = capi.make_params(p=0.01, keep_pileups=True)
    ctx = capi.Context(util.oracle_api(), satcase.CHROM_LEN, par)
    ctx.sample_begin(False, None)
    ctx.push_intervals(recs)
    ctx.sample_pileup()
    n_over, n_under, lst = ctx.sample_skipped(False)
    assert (n_over, n_under) == (meta["n_overflow"], meta["n_underflow"])
    # the skipped alignments themselves: record index -> template number (the SAM's read name) and coordinates
    tmpl = []
    for n, pl in enumerate(satcase.templates()):
        tmpl += [n] * len(pl)
    h = hashlib.sha256()
    for kind in (0, 1):
        for v in lst[(lst & np.uint64(1)) == kind]:
            i = int(v >> np.uint64(1))
            h.update(("f%d chr1 %d %d\n" % (tmpl[i], recs[i, 1], recs[i, 2])).encode())
    assert h.hexdigest() == meta["skipped_sha256"]
    st = ctx.replicate_end()
    assert abs(st.lambda_ - meta["lambda"][0]) < 1e-5
    peaks, rs = ctx.call_peaks()
    got = host.format_narrowpeak(peaks, ["chr1"])
    want = open(os.path.join(util.GOLDEN, "sat_hot.narrowPeak")).read().split("\n")[:-1]
    assert got == want and len(got) == meta["peaks"]
    assert util.sha_lines(host.format_log(ctx, ["chr1"], False, thr=par.min_pqval)) == meta["log_sha256"]
    assert util.sha_lines(host.format_pile(ctx, ["chr1"], 0)) == meta["pile_sha256"]
