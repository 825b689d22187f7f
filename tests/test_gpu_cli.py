"""The drop-in program: genrich-b200 (host C: SAM/BAM decode, pairing, fragment
inference, text writers) over the CUDA library, run as a process on the SAM view of
each case and compared with (a) the narrowPeak file the unmodified reference wrote
(tests/golden) and (b) the -f / -k / -b text the pinned oracle produces."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import util
from cases import CASES, BY_NAME
from genrich_b200 import host

pytestmark = pytest.mark.gpu

CLI = os.path.join(util.ROOT, "genrich_b200", "bin", "genrich-b200")
if os.environ.get("GR_EMU_AS_CUDA"):       # development aid: the host program over the CPU-emulated library (tests/emu)
    subprocess.check_call(["make", "-s", "-C", os.path.join(util.ROOT, "tests", "emu"), "_build/genrich-b200-emu"])
    CLI = os.path.join(util.ROOT, "tests", "emu", "_build", "genrich-b200-emu")


def _close_lines(got, want, float_cols, tol=1e-4):
    assert len(got) == len(want), (len(got), len(want))
    for g, w in zip(got, want):
        if g == w:
            continue
        gf, wf = g.split("\t"), w.split("\t")
        assert len(gf) == len(wf), (g, w)
        for i, (a, b) in enumerate(zip(gf, wf)):
            if a == b:
                continue
            assert i in float_cols, (g, w)
            fa, fb = float(a), float(b)
            assert abs(fa - fb) <= tol + 2e-6 * abs(fb) + 1.1e-6, (g, w)


def _run_cli(case, td, extra=()):
    tfiles, cfiles = util.write_case_sams(case, td)
    out, logf, pile, bed = (os.path.join(td, x) for x in ("o.np", "o.f", "o.k", "o.b"))
    cmd = [CLI, "-t", ",".join(tfiles), "-o", out, "-f", logf, "-k", pile, "-b", bed, "-v"] + case.ref_args() + list(extra)
    if any(c != "null" for c in cfiles):
        cmd += ["-c", ",".join(cfiles)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    rd = lambda p: open(p).read().split("\n")[:-1]
    return rd(out), rd(logf), rd(pile), rd(bed), r.stderr


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_cli_matches_reference_and_oracle(case, tmp_path):
    np_lines, log, pile, bed, err = _run_cli(case, str(tmp_path))
    meta, gold = util.golden(case)
    # narrowPeak vs the reference's own file
    assert len(np_lines) == len(gold) == meta["peaks"]
    for g, w in zip(np_lines, gold):
        gf, wf = g.split("\t"), w.split("\t")
        assert gf[:4] == wf[:4] and gf[5] == wf[5] and gf[9] == wf[9], (g, w)   # chrom start end name . summit
        assert abs(int(gf[4]) - int(wf[4])) <= 1
        assert abs(float(gf[6]) - float(wf[6])) <= 1e-4 * max(1.0, float(wf[6]))
        assert abs(float(gf[7]) - float(wf[7])) <= 1e-4 + 1e-6
        assert abs(float(gf[8]) - float(wf[8])) <= 1e-4 + 1e-6
    # -v scalars exactly as the reference printed them
    import re
    lam = [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)]
    fac = [float(x) for x in re.findall(r"Scaling factor for control pileup: ([0-9.]+)", err)]
    assert lam == meta["lambda"] and fac == meta["factor"]
    assert int(re.search(r"Genome length: (\d+)bp", err).group(1)) == meta["genome_len"]
    assert int(re.search(r"Peaks identified: (\d+)", err).group(1)) == meta["peaks"]
    assert int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)) == meta["peak_bp"]
    assert ("All q-values are 1" in err) == meta["all_q_one"]
    assert len(re.findall(r"prevented from extending", err)) == meta["clamp_warnings"]
    # -f / -k vs the oracle's text (itself byte-identical to the reference's, test_oracle_pin)
    ctx, res, par = util.run_case(util.oracle_api(), case)
    nrep = len(case.reps)
    want_log = util.log_lines(ctx, case, par)
    ncol = len(want_log[0].split("\t"))
    _close_lines(log, want_log, set(range(3, ncol)))
    want_pile = util.pile_lines(ctx, case)
    got_pile = [l for l in pile if not l.startswith("#")]
    _close_lines(got_pile, want_pile, {5})
    assert len(log) == meta["log_lines"] and len(got_pile) == meta["pile_lines"]
    # -b: every interval the host emitted, clamped, in order
    want = []
    for r, (e, c, _) in enumerate(util.case_inputs(case)):
        for arr, tag in ((e, "E"), (c, "C")):
            if arr is None:
                continue
            L = np.asarray(case.chrom_len, dtype=np.int64)[arr[:, 0]]
            s = np.maximum(arr[:, 1].astype(np.int64), 0)
            t = np.minimum(arr[:, 2].astype(np.int64), L)
            want += ["chr%d\t%d\t%d\t%d_%s_%d" % (a + 1, b, c2, k, tag, r) for a, b, c2, k in zip(arr[:, 0], s, t, arr[:, 3])]
    got = []
    for l in bed:
        f = l.split("\t")
        nm = f[3].split("_")
        got.append("%s\t%s\t%s\t%s_%s_%s" % (f[0], f[1], f[2], nm[-3], nm[-2], nm[-1]))
    assert got == want


_sam_to_bam = util.sam_to_bam


def test_cli_bam_and_gz_inputs(tmp_path):
    """BAM (BGZF) and gzip-compressed SAM inputs give the same peaks as plain SAM."""
    case = BY_NAME["c5_multimap_ctrl_p"]
    td = str(tmp_path)
    tfiles, cfiles = util.write_case_sams(case, td)
    base = os.path.join(td, "plain.np")
    args = case.ref_args()
    subprocess.check_call([CLI, "-t", tfiles[0], "-c", cfiles[0], "-o", base] + args)
    bam_t, bam_c = os.path.join(td, "t.bam"), os.path.join(td, "c.bam")
    _sam_to_bam(tfiles[0], bam_t)
    _sam_to_bam(cfiles[0], bam_c)
    o2 = os.path.join(td, "bam.np")
    subprocess.check_call([CLI, "-t", bam_t, "-c", bam_c, "-o", o2] + args)
    assert open(o2).read() == open(base).read()
    gz_t = os.path.join(td, "t.sam.gz")
    with open(tfiles[0], "rb") as f, gzip.open(gz_t, "wb") as g:
        g.write(f.read())
    o3 = os.path.join(td, "gz.np")
    subprocess.check_call([CLI, "-t", gz_t, "-c", cfiles[0], "-o", o3] + args)
    assert open(o3).read() == open(base).read()
    meta, gold = util.golden(case)
    assert [l.split("\t")[:3] for l in open(base).read().split("\n")[:-1]] == [l.split("\t")[:3] for l in gold]


def test_cli_errors(tmp_path):
    td = str(tmp_path)
    r = subprocess.run([CLI, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error! Need input/output files" in r.stderr
    bad = os.path.join(td, "bad.sam")
    open(bad, "w").write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:1000\n")
    r = subprocess.run([CLI, "-t", bad, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "not sorted by queryname" in r.stderr
    emp = os.path.join(td, "empty.sam")
    open(emp, "w").write("@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:chr1\tLN:1000\n")
    r = subprocess.run([CLI, "-t", emp, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Experimental sample has no analyzable fragments" in r.stderr
    r = subprocess.run([CLI, "-t", emp, "-o", os.path.join(td, "x"), "-p", "1.5"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "p-/q-value must be in (0,1]" in r.stderr


@pytest.mark.parametrize("name", ["host_r_y", "host_w", "host_r_multimap", "host_m_e"])
def test_cli_host_options(name, tmp_path):
    """-y / -w / -m / -e / -r / -R on SAM files with unpaired and discordant alignments, PCR
    duplicates and quality strings, against what the unmodified reference wrote for them
    (tests/golden/host_*): coordinates, counts and the whole -v / -R text exactly, -log10 p / q
    within 1e-4.  (All eleven host cases run byte-exact on CPU: tests/test_cli_host.py.)"""
    from hostcases import HOST_BY_NAME
    from test_cli_host import check_host_case
    out, logf, err, meta, peaks = check_host_case(CLI, HOST_BY_NAME[name], str(tmp_path), exact=False)
    assert err == meta["stderr"]
    got, gold = open(out).read().split("\n")[:-1], peaks.split("\n")[:-1]
    assert len(got) == len(gold) == meta["peaks"]
    for g, w in zip(got, gold):
        gf, wf = g.split("\t"), w.split("\t")
        assert gf[:4] == wf[:4] and gf[5] == wf[5] and gf[9] == wf[9], (g, w)
        assert abs(int(gf[4]) - int(wf[4])) <= 1
        assert abs(float(gf[6]) - float(wf[6])) <= 1e-4 * max(1.0, float(wf[6]))
        assert abs(float(gf[7]) - float(wf[7])) <= 1e-4 + 1e-6
        assert abs(float(gf[8]) - float(wf[8])) <= 1e-4 + 1e-6
    assert sum(1 for _ in open(logf)) == meta["log_lines"]


def test_cli_threaded_decode(tmp_path, monkeypatch):
    """Several host threads feeding one context (plain SAM cut at read-name boundaries; BGZF members
    inflated in parallel): the same narrowPeak file as the sequential decode, byte for byte --
    interval order does not matter to the integer pileups."""
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    monkeypatch.setenv("GB_BGZF_BATCH_BYTES", "200000")
    case = BY_NAME["c5_multimap_ctrl_p"]
    td = str(tmp_path)
    tfiles, cfiles = util.write_case_sams(case, td)
    args = case.ref_args()
    outs = []
    for th in (1, 6):
        o = os.path.join(td, "t%d.np" % th)
        subprocess.check_call([CLI, "-t", tfiles[0], "-c", cfiles[0], "-o", o, "--threads", str(th)] + args)
        outs.append(open(o).read())
    bam_t = os.path.join(td, "t.bam")
    _sam_to_bam(tfiles[0], bam_t)
    o = os.path.join(td, "bam6.np")
    subprocess.check_call([CLI, "-t", bam_t, "-c", cfiles[0], "-o", o, "--threads", "6"] + args)
    outs.append(open(o).read())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 0


@pytest.mark.parametrize("threads", [1, 4])
def test_cli_saturation_rule(threads, tmp_path, monkeypatch):
    """saveInterval 2558-2573 through the drop-in program over the CUDA library: the alignments the reference
    drops once its int16 counters saturate (tests/satcase.py: 48 466 of them, file order) are dropped on the
    device (k_sat_resolve).  narrowPeak coordinates equal the reference's file; under -v the very same
    "skipped due to overflow / underflow" lines, in the same order."""
    import hashlib
    import json
    import re
    import satcase
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    td = str(tmp_path)
    sam = os.path.join(td, "t.sam")
    satcase.write_sam(sam)
    meta = json.load(open(os.path.join(util.GOLDEN, "sat_hot.json")))
    out = os.path.join(td, "o.np")
    r = subprocess.run([CLI, "-t", sam, "-o", out, "-v", "--threads", str(threads)] + satcase.ARGS, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = open(out).read().split("\n")[:-1]
    want = open(os.path.join(util.GOLDEN, "sat_hot.narrowPeak")).read().split("\n")[:-1]
    assert len(got) == len(want) == meta["peaks"]
    _close_lines(got, want, float_cols=(6, 7, 8))
    over = re.findall(r"Warning! Read (\S+), alignment at \((\S+), (\d+)-(\d+)\) skipped due to overflow", r.stderr)
    under = re.findall(r"Warning! Read (\S+), alignment at \((\S+), (\d+)-(\d+)\) skipped due to underflow", r.stderr)
    assert (len(over), len(under)) == (meta["n_overflow"], meta["n_underflow"])
    h = hashlib.sha256()
    for a in over + under:
        h.update(("%s %s %s %s\n" % a).encode())
    assert h.hexdigest() == meta["skipped_sha256"]
    assert "Background pileup value: %f" % meta["lambda"][0] in r.stderr


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2 and not os.environ.get("GR_EMU_AS_CUDA"), reason="needs two GPUs")
@pytest.mark.parametrize("name", ["c2_ctrl_q", "c4_fisher_q", "bed_ctrl_q"])
def test_cli_sharded_over_devices(name, tmp_path):
    """genrich-b200 --gpus N on real devices (runProgram's per-chromosome loop as the multi-GPU dispatcher of the
    host program: N contexts in one process, per-chromosome sums added on the host, one p-value histogram through
    gr_bh_*_host, peaks merged in chromosome order): same narrowPeak / -f text as on one device."""
    case = BY_NAME[name]
    os.makedirs(str(tmp_path / "one"))
    os.makedirs(str(tmp_path / "many"))
    one = _run_cli(case, str(tmp_path / "one"))
    n = max(2, min(_n_gpus(), 4)) if not os.environ.get("GR_EMU_AS_CUDA") else 3
    many = _run_cli(case, str(tmp_path / "many"), extra=["--gpus", str(n)])
    assert many[0] == one[0] and len(one[0]) > 0          # narrowPeak
    assert many[1] == one[1]                              # -f
    assert [l for l in many[2] if not l.startswith("#")] == [l for l in one[2] if not l.startswith("#")]   # -k (minus the path lines)
