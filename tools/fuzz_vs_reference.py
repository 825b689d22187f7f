#!/usr/bin/env python
"""Seeded random cases against the UNMODIFIED reference binary (oracle/_ref/Genrich) -- test infrastructure, CPU only.

  python tools/fuzz_vs_reference.py cli-emu    LO HI [--gpus N]     host program over the CUDA library compiled for the
                                                                   CPU (tests/emu) vs the reference, same SAM files
  python tools/fuzz_vs_reference.py cli-oracle LO HI [--threads N] [--bam]   host program over the oracle vs the reference
  python tools/fuzz_vs_reference.py ponly      LO HI [--emu]        -P (peaks from a -f log) with random thresholds / -e / -E
  python tools/fuzz_vs_reference.py oracle     LO HI                oracle (interval view) vs the reference (SAM view)
  python tools/fuzz_vs_reference.py lib-mid    LO HI                emulated library vs oracle: 30-120 k records per sample,
                                                                   table / candidate capacities forced small (every redo path)
  python tools/fuzz_vs_reference.py sat        LO HI                random int16-saturation cases: reference == host program
                                                                   == emulated library vs oracle

LO HI = seed range.  Every output is compared byte for byte (narrowPeak, -f, -k, -b, -R, the -v text).  The cases come
from tests/fuzzcases.py; a small fixed set of seeds is part of the CPU test suite (tests/test_emu_library.py,
tests/test_cli_host.py).  Build first: `make -C oracle ref cli_twin`, `make -C tests/emu all`."""
import argparse
import gzip
import hashlib
import os
import random
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
import hostcases  # noqa: E402
from fuzzcases import random_case, random_host_options  # noqa: E402
from genrich_b200 import capi, host  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "Genrich")
TWIN = os.path.join(ROOT, "oracle", "_test", "genrich-b200-oracle")
EMU = os.environ.get("FZ_EMU", os.path.join(ROOT, "tests", "emu", "_build", "genrich-b200-emu"))
NAMES = ("exit code", "narrowPeak", "-f", "-k", "-R", "-b", "-v text")


def read_out(path):
    if not os.path.exists(path):
        return None
    b = open(path, "rb").read()
    return gzip.decompress(b) if b[:2] == b"\x1f\x8b" else b


def run_pair(case, extra, seed, other, td, gpus=0, threads=0, bam=False):
    """reference and `other` on the same (mutated) SAM / BAM / gzip files; returns the names of the outputs that differ"""
    tf, cf = util.write_case_sams(case, td)

    def conv(p, k):
        q = p.replace(".sam", ".m.sam")
        hostcases.mutate_sam(p, q, seed + k)
        if not bam:
            return q
        if (seed + k) % 2 == 0:
            b = q.replace(".sam", ".bam")
            util.sam_to_bam(q, b)
            return b
        with open(q, "rb") as fi, gzip.open(q + ".gz", "wb") as fo:
            fo.write(fi.read())
        return q + ".gz"
    tf = [conv(p, i) for i, p in enumerate(tf)]
    cf = [c if c == "null" else conv(c, 100 + i) for i, c in enumerate(cf)]
    res = []
    for exe, tag in ((REF, "A"), (other, "B")):
        f = {k: os.path.join(td, tag + "." + k) for k in ("np", "f", "k", "R", "b")}
        cmd = [exe, "-t", ",".join(tf), "-o", f["np"], "-f", f["f"], "-k", f["k"], "-b", f["b"], "-v"] + case.ref_args() + extra
        if "-r" in extra:
            cmd += ["-R", f["R"]]
        if any(c != "null" for c in cf):
            cmd += ["-c", ",".join(cf)]
        if case.bed:
            bedf = os.path.join(td, "x.bed")
            util.write_case_bed(case, bedf)
            cmd += ["-E", bedf]
        env = dict(os.environ, GB_THREAD_MIN_BYTES="1", GR_FUSED="1", GR_FUSED_MIN="1")
        if tag == "B" and gpus:
            cmd += ["--gpus", str(gpus)]
            env["EMU_DEVICES"] = str(gpus)
        if tag == "B" and threads:
            cmd += ["--threads", str(threads)]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
        res.append([r.returncode] + [read_out(p) for p in f.values()] + [r.stderr.replace(tag + ".", "X.")])
    return [n for n, a, b in zip(NAMES, res[0], res[1]) if a != b]


def mode_cli(seed, a, td):
    case = random_case(seed)
    extra = random_host_options(seed, case)
    r = random.Random(seed + 17)
    if a.mode == "cli-oracle":
        if r.random() < 0.3:
            extra += ["-L", str(r.choice([1000, 500000, 3000000000]))]
        if r.random() < 0.2:
            extra += ["-z"]
    other = EMU if a.mode == "cli-emu" else TWIN
    return run_pair(case, extra, seed, other, td, gpus=a.gpus, threads=a.threads, bam=a.bam), (case, extra)


def mode_ponly(seed, a, td):
    case = random_case(seed)
    r = random.Random(seed * 7 + 1)
    tf, cf = util.write_case_sams(case, td)
    logf = os.path.join(td, "log.f")
    cmd = [REF, "-t", ",".join(tf), "-o", os.path.join(td, "o.np"), "-f", logf] + case.ref_args()
    if any(c != "null" for c in cf):
        cmd += ["-c", ",".join(cf)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    if subprocess.run(cmd, stderr=subprocess.DEVNULL).returncode:
        return [], None
    pa = ["-q", repr(r.choice([0.01, 0.05, 0.5]))] if case.q is not None and r.random() < 0.6 else ["-p", repr(r.choice([0.001, 0.01, 0.1, 0.5]))]
    if r.random() < 0.5:
        pa += ["-a", repr(r.choice([0.0, 50.0, 500.0]))]
    if r.random() < 0.4:
        pa += ["-l", str(r.choice([0, 100, 300]))]
    if r.random() < 0.5:
        pa += ["-g", str(r.choice([0, 10, 100, 2000]))]
    if r.random() < 0.3 and len(case.chrom_len) > 1:
        pa += ["-e", "chr%d" % (r.randrange(len(case.chrom_len)) + 1)]
    if r.random() < 0.3:
        nb = os.path.join(td, "y.bed")
        with open(nb, "w") as f:
            for _ in range(r.randrange(1, 5)):
                c = r.randrange(len(case.chrom_len))
                s = r.randrange(0, max(case.chrom_len[c], 2))
                f.write("chr%d\t%d\t%d\n" % (c + 1, s, s + r.choice([1, 100, 5000, 300000])))
        pa += ["-E", nb]
    res = []
    for exe, tag in ((REF, "A"), (EMU if a.emu else TWIN, "B")):
        out = os.path.join(td, tag + ".np")
        rr = subprocess.run([exe, "-P", "-f", logf, "-o", out, "-v"] + pa, stderr=subprocess.PIPE, text=True, timeout=600)
        res.append((rr.returncode, read_out(out), rr.stderr.replace(tag + ".np", "X.np")))
    return [n for n, x, y in zip(("exit code", "narrowPeak", "-v text"), res[0], res[1]) if x != y], (case, pa)


def sha_file(path, skip_hash_lines=False):
    h = hashlib.sha256()
    for line in open(path, "rb"):
        if not (skip_hash_lines and line.startswith(b"#")):
            h.update(line)
    return h.hexdigest()


def mode_oracle(seed, a, td):
    case = random_case(seed, holes_with_multimap=False)
    tf, cf = util.write_case_sams(case, td)
    out, logf, pile = (os.path.join(td, x) for x in ("o.np", "log.f", "pile.k"))
    cmd = [REF, "-t", ",".join(tf), "-o", out, "-f", logf, "-k", pile] + case.ref_args()
    if any(c != "null" for c in cf):
        cmd += ["-c", ",".join(cf)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    rc = subprocess.run(cmd, stderr=subprocess.DEVNULL).returncode
    try:
        ctx, res, par = util.run_case(util.oracle_api(), case)
    except capi.GenrichError as e:
        return ([] if rc else ["oracle status %d, the reference ran" % e.status]), case
    if rc:
        return ["the reference failed, the oracle ran"], case
    d = []
    if host.format_narrowpeak(res.peaks, util.names_of(case)) != open(out).read().split("\n")[:-1]:
        d.append("narrowPeak")
    if util.sha_lines(util.log_lines(ctx, case, par)) != sha_file(logf):
        d.append("-f")
    if util.sha_lines(util.pile_lines(ctx, case)) != sha_file(pile, True):
        d.append("-k")
    return d, case


def emu_api():
    import test_emu_library as T
    return T, capi.Api(T.LIB, "gr_")


class Env:
    def __init__(self):
        self.keys = []

    def setenv(self, k, v):
        os.environ[k] = v
        self.keys.append(k)

    def undo(self):
        for k in self.keys:
            os.environ.pop(k, None)
        self.keys = []


def mode_lib_mid(seed, a, td):
    from cases import Case, Sample
    T, api = emu_api()
    r = np.random.RandomState(seed)
    nchrom = int(r.randint(1, 4))
    L = [int(x) for x in r.choice([200000, 600000, 1500000, 8192 * 40, 8192 * 40 + 1], nchrom)]
    reps = []
    for k in range(int(r.choice([1, 1, 2, 3]))):
        e = Sample(int(r.randint(30000, 120000)), 100 * seed + k, enrich=float(r.choice([0.1, 0.4, 0.7])),
                   spacing=int(r.choice([3000, 20000, 100000])), sigma=float(r.choice([3.0, 40.0, 150.0])), multimap=float(r.choice([0.0, 0.3])))
        c = Sample(int(r.randint(30000, 120000)), 100 * seed + 50 + k, enrich=float(r.choice([0.0, 0.1])),
                   multimap=float(r.choice([0.0, 0.3]))) if r.uniform() < 0.6 else None
        reps.append((e, c))
    use_q = r.uniform() < 0.5
    bed = []
    if r.uniform() < 0.3:
        for _ in range(int(r.randint(1, 20))):
            c = int(r.randint(nchrom))
            s = int(r.randint(0, L[c]))
            bed.append((c, s, s + int(r.choice([1, 100, 10000, 100000]))))
    case = Case("mid%d" % seed, L, reps, p=None if use_q else float(r.choice([0.01, 0.2])), q=float(r.choice([0.05, 0.5])) if use_q else None,
                min_auc=float(r.choice([0.0, 200.0])), max_gap=int(r.choice([0, 100, 5000])), atac=bool(r.uniform() < 0.3),
                atac_len=int(r.choice([100, 301])), bed=bed)
    mode = dict(T.MODES[sorted(T.MODES)[seed % len(T.MODES)]])
    if case.bed and not mode:
        mode = dict(T.FUSED)
    for knob, vals in (("GR_PAIR_CAP", [64, 1024]), ("GR_HEAD_CAP", [4, 64]), ("GR_FISHER_CAP", [64, 1024])):
        if r.uniform() < 0.5:
            mode[knob] = str(int(r.choice(vals)))
    env = Env()
    try:
        T._compare(case, api, mode, env, packed=[False, True, 6][seed % 3])
        return [], case
    except AssertionError as e:
        return ["emulated library vs oracle: " + str(e)[:300]], (case, mode)
    finally:
        env.undo()


def mode_sat(seed, a, td):
    from genrich_b200.synth import Workload
    T, api = emu_api()
    CH = [400000]
    r = np.random.RandomState(seed)
    t = []
    for h in sorted(r.choice(np.arange(20000, 380000, 1000), int(r.randint(1, 4)), replace=False).tolist()):
        kind = r.choice(["start", "end", "mixed", "half"])
        n = int(r.randint(32600, 33400))
        if kind == "start":
            t += [[(h, h + 100 + i % 300)] for i in range(n)]
        elif kind == "end":
            t += [[(h - 100 - i % 250, h)] for i in range(n)]
        elif kind == "half":
            t += [[(h, h + 120 + i % 200), (h, h + 130 + i % 170)] for i in range(n)]
            t += [[(h - 150 - i % 100, h)] for i in range(int(r.randint(0, 3000)))]
        else:
            t += [[(h, h + 100 + i % 300)] for i in range(n)]
            t += [[(h - 150 - i % 100, h)] for i in range(int(r.randint(100, 2000)))]
            h2 = h + int(r.randint(150, 400))
            t += [[(h, h2)] for i in range(int(r.randint(30000, 34000)))]
            t += [[(h2 - 120 - i % 50, h2)] for i in range(int(r.randint(0, 3000)))]
    t += [[(int(x[1]), int(x[2]))] for x in Workload(CH, 20000, seed, enrich=0.3, spacing=20000, sigma=100.0).fragments()]
    t = [t[i] for i in r.permutation(len(t))]
    recs = np.array([(0, s, e, len(pl)) for pl in t for s, e in pl], dtype=np.int32)
    sam = os.path.join(td, "t.sam")
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:chr1\tLN:%d\n" % CH[0])
        for n, pl in enumerate(t):
            for i, (s, e) in enumerate(pl):
                sec = 256 if i else 0
                r2 = max(e - 50, 0)
                f.write("f%d\t%d\tchr1\t%d\t42\t50M\t=\t%d\t%d\t*\t*\tAS:i:0\n" % (n, 99 + sec, s + 1, r2 + 1, e - s))
                f.write("f%d\t%d\tchr1\t%d\t42\t50M\t=\t%d\t%d\t*\t*\tAS:i:0\n" % (n, 147 + sec, r2 + 1, s + 1, -(e - s)))
    res = []
    for exe, tag in ((REF, "A"), (TWIN, "B")):
        f = {k: os.path.join(td, tag + "." + k) for k in ("np", "f", "k")}
        rr = subprocess.run([exe, "-t", sam, "-o", f["np"], "-f", f["f"], "-k", f["k"], "-v", "-p", "0.01", "-s", "20"],
                            stderr=subprocess.PIPE, text=True, timeout=900)
        res.append([rr.returncode] + [read_out(p) for p in f.values()] + [rr.stderr.replace(tag + ".", "X.")])
    d = [n for n, x, y in zip(("exit code", "narrowPeak", "-f", "-k", "-v text"), res[0], res[1]) if x != y]
    os.environ.update(GR_FUSED="1", GR_FUSED_MIN="1")
    outs = []
    for lib in (util.oracle_api(), api):
        ctx = capi.Context(lib, CH, capi.make_params(p=0.01, keep_pileups=True))
        ctx.sample_begin(False, None)
        ctx.push_intervals(recs)
        ctx.sample_pileup()
        sk = ctx.sample_skipped(False)
        st = ctx.replicate_end()
        pk, _ = ctx.call_peaks()
        outs.append((sk[0], sk[1], sk[2].tobytes(), pk.tobytes(), st.lambda_))
    if outs[0] != outs[1]:
        d.append("emulated library vs oracle")
    return d, "skipped alignments: %d" % res[0][4].count("skipped due to")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cli-emu", "cli-oracle", "ponly", "oracle", "lib-mid", "sat"])
    ap.add_argument("lo", type=int)
    ap.add_argument("hi", type=int)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--bam", action="store_true")
    ap.add_argument("--emu", action="store_true")
    a = ap.parse_args()
    fn = {"cli-emu": mode_cli, "cli-oracle": mode_cli, "ponly": mode_ponly, "oracle": mode_oracle, "lib-mid": mode_lib_mid, "sat": mode_sat}[a.mode]
    bad = []
    for seed in range(a.lo, a.hi):
        td = tempfile.mkdtemp()
        try:
            d, info = fn(seed, a, td)
            if d:
                bad.append((seed, d, info))
        except Exception as e:                     # a crash of the harness is a finding too
            bad.append((seed, repr(e)[:300], None))
        finally:
            shutil.rmtree(td, ignore_errors=True)
    print("%s: %d cases, %d with differences" % (a.mode, a.hi - a.lo, len(bad)))
    for b in bad:
        print("  ", b)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
