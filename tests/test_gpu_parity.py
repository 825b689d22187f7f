"""GPU parity: the CUDA path (through the C-ABI) against the pinned CPU oracle on the
same seeded inputs, against the reference's own narrowPeak files (tests/golden), and
-- at sizes the oracle cannot reach in seconds -- through size-independent
properties.  Bars: bit-exact for coordinates, interval partitions, pileup floats,
lambda, scale factor; <= 1e-4 absolute on -log10 p / -log10 q (BASELINE.json)."""
import numpy as np
import pytest

import util
from cases import CASES, BY_NAME, Case, Sample
from genrich_b200 import capi, host
from genrich_b200.synth import Workload

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def _cmp_intervals(g, o, exact_val, what):
    assert (g is None) == (o is None), what
    if g is None:
        return 0.0
    assert np.array_equal(g.end, o.end), what + ": interval ends differ"
    if exact_val:
        assert np.array_equal(_bits(g.val), _bits(o.val)), what + ": values differ"
        return 0.0
    fin = np.isfinite(o.val) & (o.val < 3e38)
    assert np.array_equal(_bits(g.val[~fin]), _bits(o.val[~fin])), what
    d = np.max(np.abs(g.val[fin].astype(np.float64) - o.val[fin])) if fin.any() else 0.0
    assert d <= TOL, (what, d)
    return d


def _parse_np(lines):
    rows = []
    for l in lines:
        f = l.split("\t")
        rows.append((f[0], int(f[1]), int(f[2]), int(f[4]), float(f[6]), float(f[7]), float(f[8]), int(f[9])))
    return rows


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_case_matches_oracle_and_reference(case):
    inputs = util.case_inputs(case)
    ctx_o, res_o, par = util.run_case(util.oracle_api(), case, inputs=inputs)
    ctx_g, res_g, _ = util.run_case(capi.load_cuda(), case, inputs=inputs)
    nrep = len(case.reps)

    # scalars of every replicate: bit-exact
    for so, sg in zip(res_o.sample_stats, res_g.sample_stats):
        assert _bits(sg.lambda_) == _bits(so.lambda_)
        assert _bits(sg.factor) == _bits(so.factor)
        assert sg.frag_len == so.frag_len and sg.ctrl_frag == so.ctrl_frag
        assert (sg.n_ctrl, sg.n_pval, sg.n_clamped, sg.genome_len) == (so.n_ctrl, so.n_pval, so.n_clamped, so.genome_len)
        assert sg.n_expt == so.n_expt

    # pileups of the last replicate, p arrays of every replicate, combined p, q
    worst = 0.0
    for ci in range(len(case.chrom_len)):
        ge, oe = ctx_g.fetch(0, 0, ci), ctx_o.fetch(0, 0, ci)
        _cmp_intervals(ge, oe, True, "expt pileup chr%d" % ci)     # incl. the read-less chromosome of bed_empty_chrom (2178-2182)
        _cmp_intervals(ctx_g.fetch(1, 0, ci), ctx_o.fetch(1, 0, ci), True, "ctrl pileup chr%d" % ci)
        for r in range(nrep + (1 if nrep > 1 else 0)):
            g, o = ctx_g.fetch(2, r, ci), ctx_o.fetch(2, r, ci)
            worst = max(worst, _cmp_intervals(g, o, False, "p rep%d chr%d" % (r, ci)))
            if g is not None and r < nrep:
                assert np.array_equal(_bits(g.expt), _bits(o.expt))
                assert np.array_equal(_bits(g.ctrl), _bits(o.ctrl))
        if case.q is not None:
            worst = max(worst, _cmp_intervals(ctx_g.fetch(3, 0, ci), ctx_o.fetch(3, 0, ci), False, "q chr%d" % ci))

    # peaks vs the oracle
    a, b = res_g.peaks, res_o.peaks
    assert len(a) == len(b)
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(a[f], b[f]), f
    if len(a):
        assert np.max(np.abs(a["pval"] - b["pval"])) <= TOL
        assert np.max(np.abs(a["qval"] - b["qval"])) <= TOL
        assert np.allclose(a["auc"], b["auc"], rtol=1e-4, atol=1e-3)
    rg, ro = res_g.run_stats, res_o.run_stats
    assert (rg.genome_len, rg.n_peaks, rg.peak_bp, rg.n_intervals, rg.all_q_one, rg.n_replicates) == \
           (ro.genome_len, ro.n_peaks, ro.peak_bp, ro.n_intervals, ro.all_q_one, ro.n_replicates)
    if case.q is not None:
        assert rg.n_distinct_p == ro.n_distinct_p

    # peaks vs the reference's own narrowPeak file: coordinates and counts bit-exact
    meta, gold = util.golden(case)
    got = _parse_np(host.format_narrowpeak(a, util.names_of(case)))
    ref = _parse_np(gold)
    assert len(got) == len(ref) == meta["peaks"]
    for x, y in zip(got, ref):
        assert x[:3] == y[:3] and x[7] == y[7], (x, y)         # chrom, start, end, summit
        assert abs(x[5] - y[5]) <= TOL + 1e-6 and abs(x[6] - y[6]) <= TOL + 1e-6
        assert abs(x[4] - y[4]) <= 1e-4 * max(1.0, abs(y[4]))
        assert abs(x[3] - y[3]) <= 1
    print("%s: worst |d(-log10 p/q)| = %.3g" % (case.name, worst))


def test_mid_size_vs_oracle():
    """20 Mbp, 1.5 M + 1.5 M templates with multimapping: tens of thousands of distinct
    p-values (multi-block radix sort in BH, table growth), many look-back tiles."""
    case = Case("mid", [12_000_000, 8_000_000],
                [(Sample(1_500_000, 61, enrich=0.3, spacing=30000, sigma=100.0, multimap=0.3),
                  Sample(1_500_000, 62, enrich=0.0, multimap=0.3))], q=0.05)
    inputs = util.case_inputs(case)
    ctx_o, res_o, par = util.run_case(util.oracle_api(), case, inputs=inputs)
    ctx_g, res_g, _ = util.run_case(capi.load_cuda(), case, inputs=inputs)
    assert _bits(res_g.sample_stats[0].lambda_) == _bits(res_o.sample_stats[0].lambda_)
    assert _bits(res_g.sample_stats[0].factor) == _bits(res_o.sample_stats[0].factor)
    worst = 0.0
    for ci in range(2):
        _cmp_intervals(ctx_g.fetch(0, 0, ci), ctx_o.fetch(0, 0, ci), True, "expt")
        _cmp_intervals(ctx_g.fetch(1, 0, ci), ctx_o.fetch(1, 0, ci), True, "ctrl")
        worst = max(worst, _cmp_intervals(ctx_g.fetch(2, 0, ci), ctx_o.fetch(2, 0, ci), False, "p"))
        worst = max(worst, _cmp_intervals(ctx_g.fetch(3, 0, ci), ctx_o.fetch(3, 0, ci), False, "q"))
    a, b = res_g.peaks, res_o.peaks
    assert len(a) == len(b) and len(a) > 100
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(a[f], b[f]), f
    assert res_g.run_stats.n_distinct_p == res_o.run_stats.n_distinct_p
    print("mid: %d peaks, %d distinct p, worst %.3g" % (len(a), res_g.run_stats.n_distinct_p, worst))


def test_repeatable_and_chunking_invariant():
    """Same input pushed in different chunkings / twice gives identical bits
    (integer atomics and the fixed-point length sums are order-free)."""
    case = BY_NAME["c5_multimap_ctrl_p"]
    inputs = util.case_inputs(case)
    api = capi.load_cuda()
    outs = []
    for chunk in (1 << 22, 7001, 1 << 22):
        ctx = capi.Context(api, case.chrom_len, util.case_params(case))
        res = host.run_replicates(ctx, inputs, chunk=chunk)
        p = [ctx.fetch(2, 0, c) for c in range(len(case.chrom_len))]
        outs.append((res, p))
    for res, p in outs[1:]:
        assert res.peaks.tobytes() == outs[0][0].peaks.tobytes()
        assert res.sample_stats[0].frag_len == outs[0][0].sample_stats[0].frag_len
        for x, y in zip(p, outs[0][1]):
            assert np.array_equal(x.end, y.end) and np.array_equal(_bits(x.val), _bits(y.val))


def test_packed_records_same_bits():
    """gr_push_packed (8-byte records) == gr_push_intervals on the same intervals, including
    records that cannot be packed (negative start, >= 16384 bp) and travel the 16-byte way."""
    api = capi.load_cuda()
    for name in ("c5_multimap_ctrl_p", "c2_ctrl_q"):
        case = BY_NAME[name]
        inputs = [list(r) for r in util.case_inputs(case)]
        extra = np.array([[0, -40, 180, 1], [0, 1000, 1000 + 20000, 2], [1 % len(case.chrom_len), 5, 16388, 1]], np.int32)
        inputs[0][0] = np.concatenate([inputs[0][0], extra])
        outs = []
        for packed in (False, True):
            ctx = capi.Context(api, case.chrom_len, util.case_params(case))
            res = host.run_replicates(ctx, inputs, chunk=50001, packed=packed)
            outs.append((res, [ctx.fetch(2, 0, c) for c in range(len(case.chrom_len))]))
        (ra, pa), (rb, pb) = outs
        assert ra.peaks.tobytes() == rb.peaks.tobytes() and len(ra.peaks) > 0
        assert ra.sample_stats[0].frag_len == rb.sample_stats[0].frag_len
        for x, y in zip(pa, pb):
            assert np.array_equal(x.end, y.end) and np.array_equal(_bits(x.val), _bits(y.val))


def test_packed6_records_same_bits(monkeypatch):
    """gr_push_packed6 (6-byte records, expanded on the device) == gr_push_intervals on the same
    intervals: every scan path, records that do not fit the format (clamped, 4096 bp or longer)
    mixed in through the 8- and 16-byte forms, a -E case, a skipped chromosome."""
    api = capi.load_cuda()
    for env in (FUSED, PLAIN):
        for k in ("GR_FUSED", "GR_FUSED_MIN", "GR_SB_MIN"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for name in ("c5_multimap_ctrl_p", "c2_ctrl_q", "bed_ctrl_q"):
            case = BY_NAME[name]
            inputs = [list(r) for r in util.case_inputs(case)]
            extra = np.array([[0, -40, 180, 1], [0, 1000, 1000 + 20000, 2], [1 % len(case.chrom_len), 5, 4200, 1],
                              [0, case.chrom_len[0] - 30, case.chrom_len[0] + 50, 3]], np.int32)
            inputs[0][0] = np.concatenate([inputs[0][0], extra])
            outs = []
            for packed in (False, 6):
                ctx = capi.Context(api, case.chrom_len, util.case_params(case))
                if case.bed:
                    ctx.set_exclusions(case.bed)
                if packed:
                    assert ctx.pack6_layout() is not None
                res = host.run_replicates(ctx, inputs, chunk=50001, packed=packed)
                outs.append((res, [ctx.fetch(2, 0, c) for c in range(len(case.chrom_len))]))
            (ra, pa), (rb, pb) = outs
            assert ra.peaks.tobytes() == rb.peaks.tobytes() and len(ra.peaks) > 0
            assert ra.sample_stats[0].frag_len == rb.sample_stats[0].frag_len
            assert ra.sample_stats[0].n_clamped == rb.sample_stats[0].n_clamped
            for x, y in zip(pa, pb):
                assert np.array_equal(x.end, y.end) and np.array_equal(_bits(x.val), _bits(y.val))
    # a cell beyond the layout is an error, not a crash
    ctx = capi.Context(api, [20000, 9000], capi.make_params(p=0.01))
    ctx.sample_begin(False)
    ctx.push_packed6(np.array([[0xFFFF, 0xFFFF, 100 | (1 << 12)]], dtype=np.uint16))
    with pytest.raises(capi.GenrichError) as e:
        ctx.sample_pileup()
    assert e.value.status == 9


def test_bucketed_build_same_bits(monkeypatch):
    """Large samples build the delta array block by block from bucketed records, small ones use
    atomic reductions: same array, same everything after it.  GR_SB_MIN forces either way."""
    api = capi.load_cuda()
    for name in ("c5_multimap_ctrl_p", "c3_atac_q", "fisher_missing_chrom"):
        case = BY_NAME[name]
        inputs = [list(r) for r in util.case_inputs(case)]
        # an interval longer than a bucket entry can hold, and one that spans three blocks
        extra = np.array([[0, 1000, 41000, 2], [0, 8000, 3 * 8192 + 5, 1], [0, 8191, 8193, 4]], np.int32)
        inputs[0][0] = np.concatenate([inputs[0][0], extra])
        outs = []
        for sb_min, packed in (("1", False), ("1000000000", False), ("1", True)):
            monkeypatch.setenv("GR_SB_MIN", sb_min)
            ctx = capi.Context(api, case.chrom_len, util.case_params(case))
            res = host.run_replicates(ctx, inputs, chunk=30011, packed=packed)
            outs.append((res, [ctx.fetch(2, 0, c) for c in range(len(case.chrom_len))]))
        monkeypatch.delenv("GR_SB_MIN")
        for res, p in outs[1:]:
            assert res.peaks.tobytes() == outs[0][0].peaks.tobytes() and len(res.peaks) > 0
            assert res.sample_stats[0].frag_len == outs[0][0].sample_stats[0].frag_len
            for x, y in zip(p, outs[0][1]):
                if x is None or y is None:
                    assert x is None and y is None
                    continue
                assert np.array_equal(x.end, y.end) and np.array_equal(_bits(x.val), _bits(y.val))


def _run_modes(monkeypatch, chrom_len, par, inputs, modes, chunk=30011, packed=False):
    """The same replicates through several scan paths (environment read by gr_create)."""
    api = capi.load_cuda()
    outs = []
    for env in modes:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = capi.Context(api, chrom_len, par)
        res = host.run_replicates(ctx, inputs, chunk=chunk, packed=packed)
        outs.append((res, [[ctx.fetch(w, 0, c) for c in range(len(chrom_len))] for w in (0, 1, 2)]))
        for k in env:
            monkeypatch.delenv(k)
    return outs


def _same_outs(a, b, what):
    (ra, pa), (rb, pb) = a, b
    assert ra.peaks.tobytes() == rb.peaks.tobytes(), what
    for sa, sb in zip(ra.sample_stats, rb.sample_stats):
        assert (sa.frag_len, sa.ctrl_frag, sa.n_expt, sa.n_ctrl, sa.n_pval, sa.n_clamped) == \
               (sb.frag_len, sb.ctrl_frag, sb.n_expt, sb.n_ctrl, sb.n_pval, sb.n_clamped), what
        assert _bits(sa.lambda_) == _bits(sb.lambda_) and _bits(sa.factor) == _bits(sb.factor), what
    for wa, wb in zip(pa, pb):
        for x, y in zip(wa, wb):
            assert (x is None) == (y is None), what
            if x is not None:
                assert np.array_equal(x.end, y.end), what
                assert np.array_equal(_bits(x.val), _bits(y.val)), what


FUSED = {"GR_FUSED": "1", "GR_FUSED_MIN": "1"}
PLAIN = {"GR_FUSED": "0", "GR_SB_MIN": "1000000000"}


def test_fused_scan_same_bits(monkeypatch):
    """The fused path (events bucketed per 8192-cell block, the block assembled and scanned in
    shared memory, no delta array in HBM) against the plain scatter + streaming scan: every
    seeded case, the edge inputs, intervals that span several blocks, packed records."""
    for case in CASES:
        if case.bed:
            continue                               # -E regions exist in the fused scan only
        inputs = [list(r) for r in util.case_inputs(case)]
        extra = np.array([[0, 1000, 41000, 2], [0, 8000, 3 * 8192 + 5, 1], [0, 8191, 8193, 4], [0, 8192, 8192, 3],
                          [0, 16383, 16384, 5], [0, 0, 8192, 6]], np.int32)
        extra = extra[extra[:, 2] <= case.chrom_len[0]]
        inputs[0][0] = np.concatenate([inputs[0][0], extra])
        par = util.case_params(case)
        outs = _run_modes(monkeypatch, case.chrom_len, par, inputs,
                          (PLAIN, FUSED, dict(FUSED, GR_FUSED_CTA="1"), dict(FUSED, GR_FUSED_RANK="1")))
        _same_outs(outs[0], outs[1], case.name)            # the fused scan, form chosen on the device == scatter + k_scan_stream
        _same_outs(outs[0], outs[2], case.name + " cta")   # k_fb_scan forced (the form -E contexts and full blocks get)
        _same_outs(outs[0], outs[3], case.name + " rank")  # k_fr_scan forced
        assert len(outs[0][0].peaks) > 0 or case.name == "null_q"
    # packed records through the fused path
    case = BY_NAME["c2_ctrl_q"]
    inputs = util.case_inputs(case)
    par = util.case_params(case)
    a = _run_modes(monkeypatch, case.chrom_len, par, inputs, (PLAIN,))[0]
    b = _run_modes(monkeypatch, case.chrom_len, par, inputs, (FUSED,), packed=True, chunk=7001)[0]
    _same_outs(a, b, "packed")
    # edge inputs: chromosome ends on block boundaries, one-base chromosome, empty interval
    L = [5000, 8192, 8191, 1, 20000, 16384, 16383]
    recs = np.array([
        [0, 0, 5000, 1], [0, -50, 10, 2], [0, 4990, 6000, 3],
        [1, 0, 1, 1], [1, 8191, 8192, 1],
        [2, 8190, 8191, 10], [2, 0, 8191, 8],
        [3, 0, 1, 1],
        [4, 100, 100, 5],
        [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6],
        [4, 19999, 25000, 4],
        [5, 0, 16384, 1], [5, 8191, 8192, 2], [5, 8192, 8193, 2], [5, 16383, 16384, 3],
        [6, 0, 16383, 1], [6, 8100, 16383, 2], [6, 16382, 16383, 3],
    ], dtype=np.int32)
    par = capi.make_params(p=0.2, min_auc=0.5, keep_pileups=True)
    modes = (PLAIN, FUSED, dict(FUSED, GR_FUSED_CTA="1"), dict(FUSED, GR_FUSED_RANK="1"))
    outs = _run_modes(monkeypatch, L, par, [(recs, None)], modes)
    for o in outs[1:]:
        _same_outs(outs[0], o, "edge")
    assert outs[1][0].sample_stats[0].n_clamped == 3


def test_fused_scan_large(monkeypatch):
    """200 Mbp / 4 M + 4 M fragments with hot spots (thousands of events in one block, more than
    two rounds of bucket entries, many pages per run owner): fused == bucketed build == exact sums."""
    L = [60_000_000, 50_000_000, 40_000_000, 30_000_000, 20_000_000]
    t = Workload(L, 4_000_000, 101, enrich=0.5, spacing=400000, sigma=60.0).fragments()
    c = Workload(L, 4_000_000, 102, enrich=0.0).fragments()
    par = capi.make_params(p=0.01, min_auc=20.0)
    modes = ({"GR_FUSED": "0"}, {"GR_FUSED": "1"}, {"GR_FUSED": "1", "GR_FUSED_CTA": "1"}, {"GR_FUSED": "1", "GR_FUSED_RANK": "1"})
    outs = _run_modes(monkeypatch, L, par, [(t, c)], modes, chunk=1 << 22)
    for o, md in zip(outs[1:], modes[1:]):
        _same_outs(outs[0], o, "large %s" % md)
    st = outs[1][0].sample_stats[0]
    assert st.frag_len == float(np.sum((t[:, 2] - t[:, 1]).astype(np.int64)))
    assert len(outs[1][0].peaks) > 100


def test_scan_form_chosen_on_the_device(monkeypatch):
    """Both scan kernels are launched and the sample picks one (form_skip): a flat sample takes the rank form, one
    whose entries sit in full blocks (>= 1024 entries per 8192 cells: a deep sample, ATAC pile-ups) the CTA form;
    the statistic is exact, and either choice gives the bits of the plain path."""
    api = capi.load_cuda()
    L = [3_000_000, 1_000_000]
    par = capi.make_params(p=0.01)
    flat = Workload(L, 60_000, 5, enrich=0.2, spacing=100000, sigma=80.0).fragments()
    deep = Workload(L, 900_000, 6, enrich=0.3, spacing=100000, sigma=80.0).fragments()
    for k, v in FUSED.items():
        monkeypatch.setenv(k, v)
    for recs, want in ((flat, 0), (deep, 1)):
        ctx = capi.Context(api, L, par)
        res = host.run_replicates(ctx, [(recs, None)], chunk=1 << 20)
        form, hot, ent = ctx.scan_form()
        # the statistic, recomputed: entries per block (an interval that crosses a block border has two)
        off = np.concatenate([[0], np.cumsum([(l + 1 + 8191) // 8192 for l in L])])[:-1]
        s = np.clip(recs[:, 1], 0, None); e = np.minimum(recs[:, 2], np.asarray(L)[recs[:, 0]])
        bs, be = off[recs[:, 0]] + s // 8192, off[recs[:, 0]] + e // 8192
        cnt = np.bincount(np.concatenate([bs, be[be != bs]]))
        assert ent == int(cnt.sum()) and hot == int(cnt[cnt >= 1024].sum())
        assert form == want == int(hot * 4 > ent)
        for k in FUSED:
            monkeypatch.delenv(k)
        ref = _run_modes(monkeypatch, L, par, [(recs, None)], (PLAIN,), chunk=1 << 20)[0][0]
        assert ref.peaks.tobytes() == res.peaks.tobytes() and len(res.peaks) > 0
        for k, v in FUSED.items():
            monkeypatch.setenv(k, v)


def test_many_chromosomes():
    """Around a thousand contigs of a few kb: look-back tiles, bitmap blocks and the per-chromosome sums with a
    chromosome boundary every few thousand cells.  Same lambda / factor bits and peaks as the oracle."""
    api = capi.load_cuda()
    rng = np.random.RandomState(3)
    for nchrom in (900, 1100):
        L = [int(x) for x in rng.randint(2000, 9000, nchrom)]
        def sample(n, seed, enrich):
            r = np.random.RandomState(seed)
            c = r.randint(0, nchrom, n)
            s = (r.uniform(size=n) * (np.asarray(L)[c] - 300)).astype(np.int64)
            hot = r.uniform(size=n) < enrich
            s[hot] = (np.asarray(L)[c[hot]] // 2 + r.randint(-40, 40, hot.sum()))
            return np.stack([c, s, s + r.randint(80, 300, n), np.ones(n, np.int64)], 1).astype(np.int32)
        t, c = sample(40000, 1, 0.3), sample(40000, 2, 0.0)
        par = capi.make_params(p=0.01)
        outs = []
        for a in (util.oracle_api(), api):
            ctx = capi.Context(a, L, par)
            res = host.run_replicates(ctx, [(t, c)], chunk=1 << 20)
            outs.append(res)
        o, g = outs
        assert _bits(o.sample_stats[0].lambda_) == _bits(g.sample_stats[0].lambda_)
        assert _bits(o.sample_stats[0].factor) == _bits(g.sample_stats[0].factor)
        assert o.sample_stats[0].frag_len == g.sample_stats[0].frag_len and o.sample_stats[0].ctrl_frag == g.sample_stats[0].ctrl_frag
        for f in ("chrom", "start", "end", "summit"):
            assert np.array_equal(o.peaks[f], g.peaks[f])
        assert len(g.peaks) > 5


def test_saturation_rule():
    """saveInterval 2558-2573: more than 32767 starts (32768 ends) on one base -- the reference drops
    intervals in arrival order; the device replays exactly that (k_sat_resolve).  Same dropped records,
    pileups and peaks as the oracle, which reproduces the unmodified reference's files for this case."""
    api = capi.load_cuda()
    assert util.check_saturation(api) > 0
    assert util.check_saturation(api, packed=True, chunk=50021) > 0


def test_async_path_and_retries(monkeypatch):
    """The no-round-trip path (counts, lambda and the scale factor stay on the device; tables and
    candidate buffers sized optimistically) gives the same bits as the synchronous one, also when
    the optimistic sizes are wrong and the stages have to be run again."""
    import torch
    from genrich_b200.dist import ShardedEngine
    api = capi.load_cuda()
    dev = torch.device("cuda", 0)
    for name in ("c2_ctrl_p", "c1_smoke", "c5_multimap_ctrl_p"):
        case = BY_NAME[name]
        inputs = util.case_inputs(case)
        ctx0, ref, par = util.run_case(api, case)
        for env in ({}, {"GR_PAIR_CAP": "64", "GR_HEAD_CAP": "3"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            eng = ShardedEngine(api, case.chrom_len, par, dev)
            for rep in inputs:
                expt, ctrl = rep[0], rep[1]
                eng.replicate(lambda c: c.push_intervals(expt),
                              (lambda c: c.push_intervals(ctrl)) if ctrl is not None else None, want_stats=False)
            peaks, rs = eng.call_peaks()
            for k in env:
                monkeypatch.delenv(k)
            assert peaks.tobytes() == ref.peaks.tobytes() and len(peaks) > 0
            assert rs.n_intervals == ref.run_stats.n_intervals
            st = eng.ctx.replicate_stats(0)
            assert st.lambda_ == ref.sample_stats[0].lambda_ and st.factor == ref.sample_stats[0].factor
            for c in range(len(case.chrom_len)):
                a, b = eng.ctx.fetch(2, 0, c), ctx0.fetch(2, 0, c)
                assert np.array_equal(a.end, b.end) and np.array_equal(_bits(a.val), _bits(b.val))


def test_fisher_table_growth(monkeypatch):
    """The chi-square tails of the Fisher combine go through a table of distinct sums (k_fsum_*); a table that
    turns out too small is grown and the stage run again: same combined p, q and peaks."""
    api = capi.load_cuda()
    case = BY_NAME["c4_fisher_q"]
    inputs = util.case_inputs(case)
    ctx0, ref, par = util.run_case(api, case, inputs=inputs)
    monkeypatch.setenv("GR_FISHER_CAP", "64")
    ctx1, got, _ = util.run_case(api, case, inputs=inputs)
    assert got.peaks.tobytes() == ref.peaks.tobytes() and len(got.peaks) > 0
    for c in range(len(case.chrom_len)):
        for which in (2, 3):
            a, b = ctx1.fetch(which, len(case.reps), c), ctx0.fetch(which, len(case.reps), c)
            assert np.array_equal(a.end, b.end) and np.array_equal(_bits(a.val), _bits(b.val))


def test_fisher_many_replicates():
    """Ten replicates (the Fisher emit kernel takes eight per launch and continues the sums in the next), one of
    them without the last chromosome: combined p, q and peaks as the oracle's."""
    reps = [(Sample(12000, 300 + r, drop_chroms=((1,) if r == 4 else ())), Sample(12000, 400, enrich=0.0) if r == 0 else None)
            for r in range(10)]
    case = Case("fisher10", [120000, 60000], reps, q=0.05)
    inputs = util.case_inputs(case)
    ctx_o, res_o, par = util.run_case(util.oracle_api(), case, inputs=inputs)
    ctx_g, res_g, _ = util.run_case(capi.load_cuda(), case, inputs=inputs)
    a, b = res_g.peaks, res_o.peaks
    assert len(a) == len(b) and len(a) > 0
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(a[f], b[f]), f
    for ci in range(2):
        for which in (2, 3):
            _cmp_intervals(ctx_g.fetch(which, 10, ci), ctx_o.fetch(which, 10, ci), False, "combined %d chr%d" % (which, ci))


def test_edge_inputs():
    api = capi.load_cuda()
    orc = util.oracle_api()
    L = [5000, 8192, 8191, 1, 20000]
    recs = np.array([
        [0, 0, 5000, 1],         # whole chromosome
        [0, -50, 10, 2],         # clamped at 0
        [0, 4990, 6000, 3],      # clamped at len
        [1, 0, 1, 1], [1, 8191, 8192, 1],   # first and last base of a block-sized chromosome
        [2, 8190, 8191, 10], [2, 0, 8191, 8],
        [3, 0, 1, 1],            # one-base chromosome
        [4, 100, 100, 5],        # empty interval: +w and -w on the same cell
        [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6],
        [4, 19999, 25000, 4],
    ], dtype=np.int32)
    for q in (None, 0.5):
        par = capi.make_params(p=0.2 if q is None else None, q=q, min_auc=0.5, keep_pileups=True)
        res = []
        for a in (orc, api):
            ctx = capi.Context(a, L, par)
            r = host.run_replicates(ctx, [(recs, None)])
            res.append((ctx, r))
        (co, ro), (cg, rg) = res
        assert ro.sample_stats[0].n_clamped == rg.sample_stats[0].n_clamped == 3
        assert _bits(ro.sample_stats[0].lambda_) == _bits(rg.sample_stats[0].lambda_)
        for ci in range(len(L)):
            _cmp_intervals(cg.fetch(0, 0, ci), co.fetch(0, 0, ci), True, "edge expt %d" % ci)
            _cmp_intervals(cg.fetch(2, 0, ci), co.fetch(2, 0, ci), False, "edge p %d" % ci)
        assert np.array_equal(rg.peaks["start"], ro.peaks["start"]) and np.array_equal(rg.peaks["end"], ro.peaks["end"])


def test_dense_breaks_vs_oracle(monkeypatch):
    """Nearly every base is an interval boundary (short fragments at ~40x): the union pass lists
    more breaks per CTA than one round of its shared-memory list holds, pages of the scan fill
    within one block, the control sweep drops and keeps boundaries side by side.  Plain and fused
    paths against the oracle, pileups / partitions bit for bit."""
    rng = np.random.default_rng(77)
    L = [70000, 9000, 8192]

    def sample(n, seed, wmax):
        r = np.random.default_rng(seed)
        ch = r.choice(len(L), size=n, p=np.asarray(L) / sum(L)).astype(np.int32)
        ln = r.integers(1, 40, size=n).astype(np.int32)
        st = (r.random(n) * (np.asarray(L)[ch] - ln)).astype(np.int32)
        w = r.choice([1, 2, 3, 4, 5, 6, 8, 10], size=n).astype(np.int32) if wmax else np.ones(n, np.int32)
        return np.stack([ch, st, st + ln, w], axis=1).astype(np.int32)
    t, c = sample(160000, 1, True), sample(120000, 2, False)
    hot = sample(30000, 3, False)
    hot = hot[hot[:, 0] == 0]
    ln = hot[:, 2] - hot[:, 1]
    hot[:, 1] = hot[:, 1] % 3000 + 30000                  # piled onto chr1:30000-33040
    hot[:, 2] = hot[:, 1] + ln
    t = np.concatenate([t, hot])
    par = capi.make_params(p=0.05, min_auc=2.0, keep_pileups=True)
    ctx_o = capi.Context(util.oracle_api(), L, par)
    res_o = host.run_replicates(ctx_o, [(t, c)])
    for env in (PLAIN, FUSED):
        for k in ("GR_FUSED", "GR_FUSED_MIN", "GR_SB_MIN"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx_g = capi.Context(capi.load_cuda(), L, par)
        res_g = host.run_replicates(ctx_g, [(t, c)])
        assert _bits(res_g.sample_stats[0].lambda_) == _bits(res_o.sample_stats[0].lambda_)
        assert _bits(res_g.sample_stats[0].factor) == _bits(res_o.sample_stats[0].factor)
        for ci in range(len(L)):
            _cmp_intervals(ctx_g.fetch(0, 0, ci), ctx_o.fetch(0, 0, ci), True, "dense expt")
            _cmp_intervals(ctx_g.fetch(1, 0, ci), ctx_o.fetch(1, 0, ci), True, "dense ctrl")
            _cmp_intervals(ctx_g.fetch(2, 0, ci), ctx_o.fetch(2, 0, ci), False, "dense p")
        n_iv = sum(len(ctx_g.fetch(2, 0, ci).end) for ci in range(len(L)))
        assert n_iv > 0.8 * sum(L)                      # the case is what it claims to be
        a, b = res_g.peaks, res_o.peaks
        assert len(a) == len(b) and len(a) > 0
        for f in ("chrom", "start", "end", "summit"):
            assert np.array_equal(a[f], b[f]), f
    del rng


def test_error_codes(monkeypatch):
    api = capi.load_cuda()
    par = capi.make_params(p=0.01)
    # start beyond the reference end -> ERRPOS (Genrich.c:2531)
    ctx = capi.Context(api, [1000], par)
    ctx.sample_begin(False)
    ctx.push_intervals(np.array([[0, 1000, 1100, 1]], dtype=np.int32))
    with pytest.raises(capi.GenrichError) as e:
        ctx.sample_pileup()
    assert e.value.status == 4
    # disallowed count -> ERRALNS (2400)
    ctx = capi.Context(api, [1000], par)
    ctx.sample_begin(False)
    ctx.push_intervals(np.array([[0, 10, 20, 7]], dtype=np.int32))
    with pytest.raises(capi.GenrichError) as e:
        ctx.sample_pileup()
    assert e.value.status == 8
    # no fragments at all -> ERREXPT (2292)
    ctx = capi.Context(api, [1000], par)
    ctx.sample_begin(False)
    with pytest.raises(capi.GenrichError) as e:
        ctx.replicate_end()
    assert e.value.status == 5
    # more starts on one base than the reference's int16 counter holds: the reference drops the 32768th
    # (saveInterval 2558-2573), and so do the oracle and the fused paths (k_sat_resolve; the full case is
    # test_saturation_rule).  The dense formulation -- a measurement aid behind GR_FUSED=0 -- only detects it.
    for env in (PLAIN, FUSED, dict(FUSED, GR_FUSED_CTA="1"), dict(FUSED, GR_FUSED_RANK="1")):
        for k in ("GR_FUSED", "GR_FUSED_MIN", "GR_SB_MIN", "GR_FUSED_CTA", "GR_FUSED_RANK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for n in (32767, 32768, 32790):
            recs = np.tile(np.array([[0, 5000, 5100, 1]], dtype=np.int32), (n, 1))
            recs[:, 2] += np.arange(n, dtype=np.int32) % 3000          # ends spread out: only the start cell is hot
            sums = []
            for a in (util.oracle_api(), api):
                ctx = capi.Context(a, [20000, 9000], par)
                ctx.sample_begin(False)
                ctx.push_intervals(recs)
                if a is api and env is PLAIN and n > 32767:
                    with pytest.raises(capi.GenrichError) as e:
                        ctx.sample_pileup()
                    assert e.value.status == 13, (env, n)
                    continue
                sums.append(ctx.sample_pileup())
                assert ctx.sample_skipped(False)[:2] == (n - 32767, 0), (env, n)
            if len(sums) == 2:
                assert np.array_equal(sums[0], sums[1])


def test_large_properties():
    """200 Mbp / 4 M + 4 M fragments: properties that need no oracle.
    sum(interval lengths) == genome; every chromosome's last end == its length;
    ends strictly increase; sum(len*val) of the experimental pileup == total
    fragment bp (exact for integer weights); peaks lie inside chromosomes, are
    ordered, and every summit is inside its peak; a second run is bit-identical."""
    L = [60_000_000, 50_000_000, 40_000_000, 30_000_000, 20_000_000]
    t = Workload(L, 4_000_000, 101, enrich=0.3, spacing=40000, sigma=100.0).fragments()
    c = Workload(L, 4_000_000, 102, enrich=0.0).fragments()
    api = capi.load_cuda()
    par = capi.make_params(p=0.01, min_auc=20.0)
    runs = []
    for _ in range(2):
        ctx = capi.Context(api, L, par)
        res = host.run_replicates(ctx, [(t, c)])
        runs.append((ctx, res))
    ctx, res = runs[0]
    st = res.sample_stats[0]
    assert st.frag_len == float(np.sum((t[:, 2] - t[:, 1]).astype(np.int64)))
    assert st.ctrl_frag == float(np.sum((c[:, 2] - c[:, 1]).astype(np.int64)))
    tot = 0
    for ci, ln in enumerate(L):
        for which in (0, 1, 2):
            iv = ctx.fetch(which, 0, ci)
            assert iv.end[-1] == ln
            assert np.all(np.diff(iv.end.astype(np.int64)) > 0)
        iv = ctx.fetch(2, 0, ci)
        tot += int(iv.end[-1])
        assert np.all(iv.val >= 0)
    assert tot == sum(L) == res.run_stats.genome_len
    pk = res.peaks
    assert len(pk) > 100
    assert np.all(pk["start"] < pk["end"]) and np.all(pk["end"] <= np.asarray(L)[pk["chrom"]])
    assert np.all(pk["summit"] < pk["end"] - pk["start"])
    key = pk["chrom"].astype(np.int64) * (1 << 32) + pk["start"]
    assert np.all(np.diff(key) > 0)
    assert np.all(pk["pval"] > par.min_pqval)
    assert runs[1][1].peaks.tobytes() == pk.tobytes()
