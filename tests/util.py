"""Shared test plumbing: build a case's interval view, run an engine, load goldens."""
from __future__ import annotations

import hashlib
import json
import os
import struct
import subprocess
import zlib

import numpy as np

from cases import Case
from genrich_b200 import capi, host
from genrich_b200.synth import Workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libgenrich_oracle.so")
REF_FUNCS = os.path.join(ORACLE_DIR, "_ref", "libref_funcs.so")

_oracle = None


def oracle_api() -> capi.Api:
    """The CPU oracle (test infrastructure), built on demand with gcc."""
    global _oracle
    if _oracle is None:
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libgenrich_oracle.so"])
        _oracle = capi.Api(ORACLE_LIB, "orc_")
    return _oracle


def names_of(case: Case):
    return ["chr%d" % (i + 1) for i in range(len(case.chrom_len))]


def sample_intervals(case: Case, s) -> np.ndarray:
    w = Workload(case.chrom_len, s.nfrag, s.seed, enrich=s.enrich, spacing=s.spacing, sigma=s.sigma,
                 multimap=s.multimap, mmax=s.mmax)
    fr = w.fragments()
    if s.drop_chroms or s.empty_chroms:
        fr = fr[~np.isin(fr[:, 0], list(s.drop_chroms) + list(s.empty_chroms))]
    return host.fragments_to_intervals(fr, atac=case.atac, atac_len=case.atac_len)


def case_inputs(case: Case):
    reps = []
    for e, c in case.reps:
        save = None
        if e.drop_chroms:
            save = np.ones(len(case.chrom_len), dtype=np.uint8)
            save[list(e.drop_chroms)] = 0
        reps.append((sample_intervals(case, e), None if c is None else sample_intervals(case, c), save))
    return reps


def case_params(case: Case, keep=True):
    return capi.make_params(p=case.p, q=case.q, min_auc=case.min_auc, min_len=case.min_len,
                            max_gap=case.max_gap, keep_pileups=keep)


def run_case(api: capi.Api, case: Case, keep=True, device=0, inputs=None):
    par = case_params(case, keep)
    ctx = capi.Context(api, case.chrom_len, par, device=device)
    if case.bed:
        ctx.set_exclusions(case.bed)
    res = host.run_replicates(ctx, inputs if inputs is not None else case_inputs(case))
    return ctx, res, par


def log_lines(ctx, case: Case, par):
    n = len(case.reps)
    if n == 1:
        return host.format_log(ctx, names_of(case), case.q is not None, thr=par.min_pqval)
    return host.format_log_multi(ctx, names_of(case), n, case.q is not None, thr=par.min_pqval)


def pile_lines(ctx, case: Case):
    out = []
    for r in range(len(case.reps)):
        out += host.format_pile(ctx, names_of(case), r)
    return out


def sha_lines(lines) -> str:
    h = hashlib.sha256()
    for l in lines:
        h.update(l.encode() + b"\n")
    return h.hexdigest()


def golden(case: Case):
    with open(os.path.join(GOLDEN, case.name + ".json")) as f:
        meta = json.load(f)
    with open(os.path.join(GOLDEN, case.name + ".narrowPeak")) as f:
        np_lines = f.read().split("\n")[:-1]
    return meta, np_lines


def write_sample_sam(case: Case, s, path: str) -> None:
    """SAM view of one sample (what the reference binary / the CLI read)."""
    w = Workload(case.chrom_len, s.nfrag, s.seed, enrich=s.enrich, spacing=s.spacing, sigma=s.sigma,
                 multimap=s.multimap, mmax=s.mmax)
    w.write_sam(path)
    if s.drop_chroms or s.empty_chroms:
        drop = {"chr%d" % (c + 1) for c in s.drop_chroms}
        empty = {"chr%d" % (c + 1) for c in s.empty_chroms}
        keep = []
        with open(path) as f:
            for line in f:
                if line.startswith("@SQ"):
                    if line.split("\t")[1][3:] in drop:
                        continue
                elif not line.startswith("@") and line.split("\t")[2] in drop | empty:
                    continue
                keep.append(line)
        with open(path, "w") as f:
            f.writelines(keep)


def write_case_bed(case: Case, path: str) -> None:
    """The -E file of a case: its regions as BED records, in the order given (unsorted, overlapping)."""
    with open(path, "w") as f:
        for c, a, b in case.bed:
            f.write("chr%d\t%d\t%d\n" % (c + 1, a, b))


def write_case_sams(case: Case, td: str):
    tfiles, cfiles = [], []
    for r, (e, c) in enumerate(case.reps):
        tp = os.path.join(td, "t%d.sam" % r)
        write_sample_sam(case, e, tp)
        tfiles.append(tp)
        if c is None:
            cfiles.append("null")
        else:
            cp = os.path.join(td, "c%d.sam" % r)
            write_sample_sam(case, c, cp)
            cfiles.append(cp)
    return tfiles, cfiles


def sam_to_bam(sam_path, bam_path):
    """Minimal BAM writer (BGZF blocks via zlib) for the tests; QUAL strings are kept."""
    refs, recs, text = [], [], []
    for line in open(sam_path):
        if line.startswith("@"):
            text.append(line)
            if line.startswith("@SQ"):
                f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                refs.append((f["SN"], int(f["LN"])))
            continue
        recs.append(line.rstrip("\n").split("\t"))
    rid = {n: i for i, (n, _) in enumerate(refs)}
    txt = "".join(text).encode()
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(txt)) + txt + struct.pack("<i", len(refs)))
    for n, l in refs:
        out += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    ops = "MIDNSHP=X"
    for f in recs:
        qn = f[0].encode() + b"\0"
        cig = []
        num = ""
        for ch in f[5]:
            if ch.isdigit():
                num += ch
            else:
                cig.append((int(num) << 4) | ops.index(ch))
                num = ""
        seq_len = sum(c >> 4 for c in cig if (c & 15) in (0, 1, 4, 7, 8))
        aux = b""
        for t in f[11:]:
            tag, ty, val = t.split(":", 2)
            if ty == "i":
                aux += tag.encode() + b"c" + struct.pack("<b", int(val))
        body = struct.pack("<iiIIiiii", rid[f[2]], int(f[3]) - 1, (4680 << 16) | (int(f[4]) << 8) | len(qn),
                           (int(f[1]) << 16) | len(cig), seq_len, rid[f[2]], int(f[7]) - 1, int(f[8]))
        qual = b"\xff" * seq_len if f[10] == "*" else bytes(ord(ch) - 33 for ch in f[10])
        assert len(qual) == seq_len
        body += qn + b"".join(struct.pack("<I", c) for c in cig) + b"\0" * ((seq_len + 1) // 2) + qual + aux
        out += struct.pack("<i", len(body)) + body
    with open(bam_path, "wb") as g:
        for i in range(0, len(out), 60000):                 # BGZF members
            chunk = bytes(out[i:i + 60000])
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            g.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
                    struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        g.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def check_saturation(api, packed=False, chunk=None):
    """The int16 saturation case (tests/satcase.py) through `api` against the pinned oracle: the very same
    records dropped (saveInterval 2558-2573, arrival order), pileups bit for bit, the same peaks."""
    import satcase
    recs = satcase.records()
    par = capi.make_params(p=0.01, keep_pileups=True)
    outs = []
    for a in (oracle_api(), api):
        ctx = capi.Context(a, satcase.CHROM_LEN, par)
        ctx.sample_begin(False, None)
        step = chunk or len(recs)
        for i in range(0, len(recs), step):
            part = recs[i:i + step]
            if packed and a is api:
                pk, rest = host.pack_records(part)
                assert not len(rest)
                ctx.push_packed(pk)
            else:
                ctx.push_intervals(part)
        ctx.sample_pileup()
        n_over, n_under, lst = ctx.sample_skipped(False)
        st = ctx.replicate_end()
        peaks, rs = ctx.call_peaks()
        outs.append((n_over, n_under, lst.copy(), st, peaks.copy(), ctx.fetch(0, 0, 0), ctx.fetch(2, 0, 0)))
    o, g = outs
    assert (g[0], g[1]) == (o[0], o[1]) and o[0] > 30000 and o[1] > 5000
    assert np.array_equal(g[2], o[2])
    assert g[3].frag_len == o[3].frag_len and np.float32(g[3].lambda_).view(np.uint32) == np.float32(o[3].lambda_).view(np.uint32)
    assert (g[3].n_expt, g[3].n_pval) == (o[3].n_expt, o[3].n_pval)
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(g[4][f], o[4][f]), f
    assert np.max(np.abs(g[4]["pval"].astype(np.float64) - o[4]["pval"])) <= 1e-4
    assert np.array_equal(g[5].end, o[5].end) and np.array_equal(g[5].val.view(np.uint32), o[5].val.view(np.uint32))
    assert np.array_equal(g[6].end, o[6].end)
    return len(g[4])
