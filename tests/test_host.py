"""Host-side logic that needs no GPU: fragment -> interval transforms, formatting,
threshold conversion, the generator's SAM/record agreement."""
import os
import subprocess

import numpy as np

import util
from cases import BY_NAME
from genrich_b200 import capi, host
from genrich_b200.synth import Workload, BIN


def _u32(x):
    return int(x) & 0xFFFFFFFF


def _i32(x):
    x = int(x) & 0xFFFFFFFF
    return x - (1 << 32) if x >= (1 << 31) else x


def _atac_loop(frags, len5, len3, adj=True):
    """saveFragAtac (Genrich.c:2728-2749) restated record by record (uint32 arithmetic)."""
    out = []
    for c, s, e, k in frags:
        s, e = _u32(s), _u32(e)
        if adj:
            s = _u32(s + 5)
            e = _u32(e - 5)
        if _u32(s + len3) >= _u32(e - len3):          # 2737: a uint32 against an int -> both as uint32 (checked against the reference binary with -d 501)
            out.append((c, _i32(s - len5), e + len5, k))
        else:
            out.append((c, _i32(s - len5), s + len3, k))
            out.append((c, _i32(e - len3), e + len5, k))
    return np.array(out, dtype=np.int32)


def test_atac_transform_matches_record_loop():
    rng = np.random.default_rng(3)
    n = 5000
    s = rng.integers(0, 100000, n)
    ln = rng.integers(50, 400, n)
    fr = np.stack([rng.integers(0, 3, n), s, s + ln, rng.choice([1, 2, 3, 4, 5, 6, 8, 10], n)], 1).astype(np.int32)
    for d in (100, 101, 37, 400):
        l5, l3 = host.atac_lengths(d)
        assert l5 + l3 == d and l3 - l5 in (0, 1)
        got = host.fragments_to_intervals(fr, atac=True, atac_len=d)
        assert np.array_equal(got, _atac_loop(fr, l5, l3))
    got = host.fragments_to_intervals(fr, atac=True, atac_len=100, atac_adj=False)
    assert np.array_equal(got, _atac_loop(fr, 50, 50, adj=False))


def test_threshold_conversion_matches_libm():
    p = capi.make_params(p=0.01)
    assert np.float32(p.min_pqval) == np.float32(2.0) and p.qval_opt == 0
    q = capi.make_params(q=0.05)
    assert abs(q.min_pqval - 1.30103) < 1e-5 and q.qval_opt == 1


def test_peak_score_and_format():
    pk = np.zeros(2, dtype=capi.PEAK_DTYPE)
    pk[0] = (0, 392, 24610, 25371, np.float32(3602.903809), np.float32(10.161201), np.float32(-1.0), 0)
    pk[1] = (1, 5, 10, 20, np.float32(0.004), np.float32(2.5), np.float32(1.75), 0)
    lines = host.format_narrowpeak(pk, ["chr1", "chrX"])
    assert lines[0] == "chr1\t24610\t25371\tpeak_0\t1000\t.\t3602.903809\t10.161201\t-1\t392"
    assert lines[1] == "chrX\t10\t20\tpeak_1\t0\t.\t0.004000\t2.500000\t1.750000\t5"


def test_generator_sam_and_records_agree(tmp_path):
    """Every SAM template written by the generator maps to the records synth_fragments emits."""
    w = Workload([50000, 30000], 2000, 77, enrich=0.3, spacing=10000, sigma=50.0, multimap=0.4, mmax=12)
    fr = w.fragments()
    sam = os.path.join(tmp_path, "x.sam")
    w.write_sam(sam)
    want = {}
    for line in open(sam):
        if line[0] == "@":
            continue
        f = line.split("\t")
        if int(f[1]) & 0x40:                       # R1 line: fragment = [pos-1, pnext-1 + 50)
            want.setdefault(f[0], []).append((int(f[2][3:]) - 1, int(f[3]) - 1, int(f[7]) - 1 + 50))
    rows = iter(fr)
    for t in range(2000):
        pl = want["f%d" % t]
        k = len(pl)
        kept = 10 if k > 10 else (k - 1 if k in (7, 9) else k)
        for i in range(kept):
            c, s, e, cnt = next(rows)
            assert (c, s, e) == pl[i] and cnt == kept
    assert next(rows, None) is None


def test_lpt_owner_covers_everything():
    L = [100, 90, 80, 10, 5, 1]
    own = host.lpt_shard(L, 3)
    assert set(own) == {0, 1, 2} and len(own) == len(L)


def test_pack_records_roundtrip():
    rng = np.random.default_rng(5)
    n = 5000
    recs = np.stack([rng.integers(0, 30, n), rng.integers(0, 2**31 - 20000, n), np.zeros(n, np.int64),
                     rng.choice([1, 2, 3, 4, 5, 6, 8, 10], n)], axis=1)
    recs[:, 2] = recs[:, 1] + rng.integers(0, 3000, n)
    recs = recs.astype(np.int32)
    recs[7] = (3, -5, 100, 1)                 # negative start
    recs[8] = (3, 100, 100 + host.PACK_MAX_LEN, 2)   # too long
    recs[9] = (3, 100, 100 + host.PACK_MAX_LEN - 1, 2)   # just fits
    packed, rest = host.pack_records(recs)
    assert len(packed) == n - 2 and len(rest) == 2
    assert rest.tolist() == [recs[7].tolist(), recs[8].tolist()]
    s = (packed & np.uint64(0xffffffff)).astype(np.int64)
    ln = ((packed >> np.uint64(32)) & np.uint64(0x3fff)).astype(np.int64)
    c = ((packed >> np.uint64(46)) & np.uint64(0x3fff)).astype(np.int64)
    k = (packed >> np.uint64(60)).astype(np.int64)
    keep = np.ones(n, bool); keep[[7, 8]] = False
    back = np.stack([c, s, s + ln, k], axis=1).astype(np.int32)
    assert np.array_equal(back, recs[keep])


def test_oracle_restates_int16_saturation():
    """More than 32767 starts on one base: the reference skips further intervals there in arrival
    order (saveInterval, Genrich.c:2558-2573), and so does the oracle: 32768 identical fragments, the last one
    dropped for overflow (what the unmodified reference prints under -v for this input), nothing dropped for 32767."""
    from genrich_b200 import capi
    import util
    par = capi.make_params(p=0.01)
    for n, want in ((32767, (0, 0, [])), (32768, (1, 0, [32767 << 1]))):
        recs = np.tile(np.array([[0, 500, 600, 1]], dtype=np.int32), (n, 1))
        ctx = capi.Context(util.oracle_api(), [2000], par)
        ctx.sample_begin(False)
        ctx.push_intervals(recs)
        ctx.sample_pileup()
        n_over, n_under, lst = ctx.sample_skipped(False)
        assert (n_over, n_under, [int(v) for v in lst]) == want


def test_pack6_records_roundtrip():
    """GR_PACK6 (6-byte records: cell of the start in the context's layout, length, count) packs what
    it can and leaves the rest; unpacking with the layout gives the records back."""
    L = [300000, 200000, 100000]
    off = np.array([0, 303104, np.uint64(0xFFFFFFFFFFFFFFFF)], dtype=np.uint64)      # chr3 not held by this context
    recs = np.array([[0, 10, 260, 1], [1, 199900, 200000, 10], [0, 299000, 300100, 2],   # past the end: clamped elsewhere
                     [0, -5, 100, 1], [1, 5, 5000, 3], [2, 10, 20, 1], [0, 0, 0, 8], [1, 0, 4095, 6]], dtype=np.int32)
    p6, rest = host.pack6_records(recs, off, L)
    assert p6.shape == (4, 3) and p6.dtype == np.uint16
    assert rest.tolist() == [[0, 299000, 300100, 2], [0, -5, 100, 1], [1, 5, 5000, 3], [2, 10, 20, 1]]
    cell = p6[:, 0].astype(np.uint64) | (p6[:, 1].astype(np.uint64) << np.uint64(16))
    ln, cnt = p6[:, 2] & 0xFFF, p6[:, 2] >> 12
    chrom = np.where(cell >= off[1], 1, 0)
    start = cell - off[chrom]
    back = np.stack([chrom, start, start + ln, cnt], axis=1).astype(np.int32)
    assert back.tolist() == [[0, 10, 260, 1], [1, 199900, 200000, 10], [0, 0, 0, 8], [1, 0, 4095, 6]]
