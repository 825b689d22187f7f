"""Host-side mirror of the reference's driver around the hot path.

``run_replicates`` is runProgram's sample loop (Genrich.c:5460-5585) followed by
findPeaks (5605) expressed over an engine :class:`~genrich_b200.capi.Context`;
``fragments_to_intervals`` is saveFragment / saveFragAtac (2754, 2728);
``format_narrowpeak`` is printPeak (885); ``format_log`` is the single-replicate
branch of printInterval (770) as logIntervals/callPeaks emit it.  All of it is
host code: the engine behind ``Context`` is the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .capi import Context, GrParams, Api

ATACADJF = 5    # Genrich.h:35
ATACADJR = -5   # Genrich.h:36


def atac_lengths(atac_len: int = 100):
    """getArgs 5796-5797: split -d into 5' and 3' parts."""
    len3 = int(np.float32(atac_len) / np.float32(2.0) + np.float32(0.5))
    return atac_len // 2, len3


def fragments_to_intervals(frags: np.ndarray, atac: bool = False, atac_len: int = 100,
                           atac_adj: bool = True) -> np.ndarray:
    """(chrom,start,end,count) fragments -> interval records for the engine.

    Non-ATAC: identity (saveFragment 2771).  ATAC (-j): one interval per cut site,
    or one merged interval when the two would touch (saveFragAtac 2733-2748).
    Clamping to [0,len] is the engine's job (saveInterval 2522-2544)."""
    frags = np.ascontiguousarray(frags, dtype=np.int32).reshape(-1, 4)
    if not atac:
        return frags
    len5, len3 = atac_lengths(atac_len)
    c = frags[:, 0].astype(np.int64)
    s = frags[:, 1].astype(np.int64)
    e = frags[:, 2].astype(np.int64)
    k = frags[:, 3].astype(np.int64)
    if atac_adj:
        s = s + ATACADJF
        e = e + ATACADJR
    # saveFragAtac 2737: `start + atacLen3 >= (signed) (end - atacLen3)` compares a uint32 with an int, i.e. both
    # as uint32 -- a fragment that ends within atacLen3 of the chromosome start wraps around and is saved as TWO
    # intervals although they overlap
    one = ((s + len3) & 0xffffffff) >= ((e - len3) & 0xffffffff)
    n1 = int(one.sum())
    n2 = len(frags) - n1
    out = np.empty((n1 + 2 * n2, 4), dtype=np.int64)
    # keep the reference's emission order: per fragment, 5' interval then 3' interval
    cnt = np.where(one, 1, 2)
    pos = np.concatenate(([0], np.cumsum(cnt)[:-1]))
    i1 = pos[one]
    out[i1, 0] = c[one]; out[i1, 1] = s[one] - len5; out[i1, 2] = e[one] + len5; out[i1, 3] = k[one]
    two = ~one
    i2 = pos[two]
    out[i2, 0] = c[two]; out[i2, 1] = s[two] - len5; out[i2, 2] = s[two] + len3; out[i2, 3] = k[two]
    out[i2 + 1, 0] = c[two]; out[i2 + 1, 1] = e[two] - len3; out[i2 + 1, 2] = e[two] + len5; out[i2 + 1, 3] = k[two]
    return out.astype(np.int32)


@dataclass
class RunResult:
    peaks: np.ndarray
    run_stats: object
    sample_stats: list = field(default_factory=list)


def run_replicates(ctx: Context, replicates, chunk: int = 1 << 22, packed: bool = False) -> RunResult:
    """replicates: list of (expt_intervals, ctrl_intervals_or_None[, save_mask]).
    packed: True -- send the records in the 8-byte GR_PACK form (what does not fit goes the 16-byte
    way); 6 -- in the 6-byte GR_PACK6 form first, then 8, then 16."""
    layout = ctx.pack6_layout() if packed == 6 else None

    def push(recs):
        if not packed:
            ctx.push_intervals(recs)
            return
        if layout is not None:                       # packed == 6: 6-byte records first, 8-byte for what is left
            p6, recs = pack6_records(recs, layout, ctx.chrom_len)
            if len(p6):
                ctx.push_packed6(p6)
            if not len(recs):
                return
        pk, rest = pack_records(recs)
        if len(pk):
            ctx.push_packed(pk)
        if len(rest):
            ctx.push_intervals(rest)
    stats = []
    for rep in replicates:
        expt, ctrl = rep[0], rep[1]
        save = rep[2] if len(rep) > 2 else None
        ctx.sample_begin(False, save)
        for i in range(0, len(expt), chunk):
            push(expt[i:i + chunk])
        if ctrl is not None:
            ctx.sample_pileup()
            ctx.sample_begin(True)
            for i in range(0, len(ctrl), chunk):
                push(ctrl[i:i + chunk])
        stats.append(ctx.replicate_end())
    peaks, rs = ctx.call_peaks()
    return RunResult(peaks, rs, stats)


def peak_score(auc, start, end) -> int:
    """printPeak 891-892 in float: MIN((unsigned)(1000*signal/(end-start)+0.5), 1000)."""
    v = np.float32(1000.0) * np.float32(auc) / np.float32(end - start) + np.float32(0.5)
    return min(int(v), 1000)


def format_narrowpeak(peaks: np.ndarray, names) -> list[str]:
    """ENCODE narrowPeak lines exactly as printPeak (Genrich.c:885-909) prints them."""
    out = []
    for i, p in enumerate(peaks):
        q = "-1" if p["qval"] == np.float32(-1.0) else "%f" % p["qval"]
        out.append("%s\t%d\t%d\tpeak_%d\t%d\t.\t%f\t%f\t%s\t%d" % (
            names[p["chrom"]], p["start"], p["end"], i,
            peak_score(p["auc"], int(p["start"]), int(p["end"])), p["auc"], p["pval"], q, p["summit"]))
    return out


def format_log(ctx: Context, names, qval: bool, thr: float | None = None) -> list[str]:
    """Single-replicate ``-f`` body (printLogHeader 699-715, printInterval 770-803).

    thr: significance threshold (adds the ``signif`` column as callPeaks does)."""
    hdr = "chr\tstart\tend\texperimental\tcontrol\t-log(p)"
    if qval:
        hdr += "\t-log(q)"
    if thr is not None:
        hdr += "\tsignif"
    out = [hdr]
    for ci in range(ctx.nchrom):
        pv = ctx.fetch(2, 0, ci)
        if pv is None:
            continue
        qv = ctx.fetch(3, 0, ci) if qval else None
        start = 0
        for m in range(len(pv.end)):
            e, c = pv.expt[m], pv.ctrl[m]
            if c == np.float32(-1.0):
                line = "%s\t%d\t%d\t%f\t%f\tNA" % (names[ci], start, pv.end[m], e, 0.0)
                if qval:
                    line += "\tNA"
            else:
                line = "%s\t%d\t%d\t%f\t%f\t%f" % (names[ci], start, pv.end[m], e, c, pv.val[m])
                if qval:
                    line += "\t%f" % qv.val[m]
                if thr is not None:
                    v = qv.val[m] if qval else pv.val[m]
                    if v > np.float32(thr):
                        line += "\t*"
            out.append(line)
            start = int(pv.end[m])
    return out


PACK_MAX_LEN = 1 << 14      # include/genrich_cuda.h GR_PACK_MAX_LEN / GR_PACK_MAX_CHROM
PACK_MAX_CHROM = 1 << 14


def pack_records(recs: np.ndarray):
    """(chrom, start, end, count) int32 records -> (uint64 GR_PACK records, the records that do not fit).

    The 8-byte form is what travels best over PCIe; whatever it cannot express (start < 0,
    an interval of 16384 bp or more, chromosome index >= 16384) stays in the 16-byte form and
    is pushed through gr_push_intervals -- the two may be mixed within a sample."""
    recs = np.ascontiguousarray(recs, dtype=np.int32).reshape(-1, 4)
    c, s, e, k = (recs[:, i].astype(np.int64) for i in range(4))
    ln = e - s
    ok = (s >= 0) & (ln >= 0) & (ln < PACK_MAX_LEN) & (c >= 0) & (c < PACK_MAX_CHROM) & (k >= 0) & (k < 16)
    if ok.all():
        sel = slice(None)
        rest = recs[:0]
    else:
        sel = ok
        rest = recs[~ok]
    packed = (s[sel].astype(np.uint64) | (ln[sel].astype(np.uint64) << np.uint64(32))
              | (c[sel].astype(np.uint64) << np.uint64(46)) | (k[sel].astype(np.uint64) << np.uint64(60)))
    return np.ascontiguousarray(packed, dtype=np.uint64), np.ascontiguousarray(rest)


PACK6_MAX_LEN = 1 << 12


def pack6_records(recs: np.ndarray, cell_offset: np.ndarray, chrom_len):
    """(chrom, start, end, count) int32 records -> (uint16 (n, 3) GR_PACK6 records, the records that
    do not fit).  cell_offset: Context.pack6_layout().  A record fits when it lies inside a
    chromosome the context holds (no clamping needed) and is shorter than 4096 bp; the rest goes on
    through pack_records / push_intervals."""
    recs = np.ascontiguousarray(recs, dtype=np.int32).reshape(-1, 4)
    c, s, e, k = (recs[:, i].astype(np.int64) for i in range(4))
    nchrom = len(cell_offset)
    cc = np.clip(c, 0, nchrom - 1)
    off = np.asarray(cell_offset, dtype=np.uint64)[cc]
    ln = e - s
    ok = ((c >= 0) & (c < nchrom) & (off != np.uint64(0xFFFFFFFFFFFFFFFF)) & (s >= 0) & (ln >= 0) & (ln < PACK6_MAX_LEN)
          & (e <= np.asarray(chrom_len, dtype=np.int64)[cc]) & (k >= 0) & (k < 16))
    sel = slice(None) if ok.all() else ok
    rest = recs[:0] if ok.all() else recs[~ok]
    cell = off[sel] + s[sel].astype(np.uint64)
    out = np.empty((cell.shape[0], 3), dtype=np.uint16)
    out[:, 0] = (cell & np.uint64(0xFFFF)).astype(np.uint16)
    out[:, 1] = ((cell >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint16)
    out[:, 2] = (ln[sel] | (k[sel] << 12)).astype(np.uint16)
    return out, np.ascontiguousarray(rest)


def lpt_shard(chrom_len, world: int) -> np.ndarray:
    """Greedy longest-processing-time assignment of chromosomes to ranks
    (the multi-GPU dispatcher that replaces runProgram's per-chromosome loop)."""
    order = np.argsort(-np.asarray(chrom_len, dtype=np.int64), kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(len(chrom_len), dtype=np.int32)
    for c in order:
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += int(chrom_len[c])
    return owner


def format_log_multi(ctx: Context, names, nrep: int, qval: bool, thr: float | None = None) -> list[str]:
    """Multi-replicate ``-f`` body (printLogHeader 676-698, printIntervalN 724-763,
    index update printLog 826-829)."""
    hdr = "chr\tstart\tend" + "".join("\t-log(p)_%d" % i for i in range(nrep)) + "\t-log(p)_comb"
    if qval:
        hdr += "\t-log(q)"
    if thr is not None:
        hdr += "\tsignif"
    out = [hdr]
    skip = np.float32(-1.0)
    for ci in range(ctx.nchrom):
        comb = ctx.fetch(2, nrep, ci)
        if comb is None:
            continue
        reps = [ctx.fetch(2, r, ci) for r in range(nrep)]
        qv = ctx.fetch(3, 0, ci) if qval else None
        idx = [0] * nrep
        start = 0
        for m in range(len(comb.end)):
            line = "%s\t%d\t%d" % (names[ci], start, comb.end[m])
            for r in range(nrep):
                if reps[r] is None or reps[r].val[idx[r]] == skip:
                    line += "\tNA"
                else:
                    line += "\t%f" % reps[r].val[idx[r]]
            if comb.val[m] == skip:
                line += "\tNA" + ("\tNA" if qval else "")
            else:
                line += "\t%f" % comb.val[m]
                if qval:
                    line += "\t%f" % qv.val[m]
            if thr is not None:
                v = qv.val[m] if qval else comb.val[m]
                if v > np.float32(thr):
                    line += "\t*"
            out.append(line)
            for r in range(nrep):
                if reps[r] is not None and reps[r].end[idx[r]] == comb.end[m]:
                    idx[r] += 1
            start = int(comb.end[m])
    return out


def format_pile(ctx: Context, names, replicate: int) -> list[str]:
    """``-k`` body for one replicate (printPileHeader 1685, printPile 1697-1715),
    without the '# experimental file' comment line."""
    out = ["chr\tstart\tend\texperimental\tcontrol\t-log(p)"]
    for ci in range(ctx.nchrom):
        pv = ctx.fetch(2, replicate, ci)
        if pv is None:
            continue
        start = 0
        for m in range(len(pv.end)):
            if pv.ctrl[m] == np.float32(-1.0):
                out.append("%s\t%d\t%d\t%f\t%f\tNA" % (names[ci], start, pv.end[m], pv.expt[m], 0.0))
            else:
                out.append("%s\t%d\t%d\t%f\t%f\t%f" % (names[ci], start, pv.end[m], pv.expt[m], pv.ctrl[m], pv.val[m]))
            start = int(pv.end[m])
    return out
