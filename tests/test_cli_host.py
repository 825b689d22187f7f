"""The host program's own C sources (genrich_b200/cli: option parsing, SAM / gz / BAM decode,
mate pairing, multimap weighting, interval transforms, narrowPeak / -f / -k writers, -v text)
checked WITHOUT a GPU: `make -C oracle cli_twin` links them against the CPU oracle instead of
libgenrich_cuda.so (oracle/cli_twin.c, test infrastructure).  Both sides of the C-ABI then use
glibc's libm like the reference, so the comparison with the files the unmodified reference wrote
(tests/golden) is BYTE FOR BYTE: every narrowPeak column, the sha256 of the whole -f and -k text,
the -v scalars.  The same program over the CUDA library is tests/test_gpu_cli.py."""
import hashlib
import os
import re
import subprocess

import pytest

import util
from cases import CASES, BY_NAME

TWIN = os.path.join(util.ORACLE_DIR, "_test", "genrich-b200-oracle")


@pytest.fixture(scope="module")
def twin():
    if os.environ.get("GR_EMU_AS_CUDA"):
        # development aid: the same tests with the host program over the CPU-emulated CUDA library (tests/emu)
        # instead of the oracle -- e.g. --gpus N over N emulated devices (set EMU_DEVICES=4)
        emu = os.path.join(util.ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", emu, "_build/genrich-b200-emu"])
        return os.path.join(emu, "_build", "genrich-b200-emu")
    subprocess.check_call(["make", "-s", "-C", util.ORACLE_DIR, "cli_twin"])
    return TWIN


def _sha(path, skip_hash_lines=False):
    h = hashlib.sha256()
    n = 0
    with open(path, "rb") as f:
        for line in f:
            if skip_hash_lines and line.startswith(b"#"):
                continue
            h.update(line)
            n += 1
    return h.hexdigest(), n


def run_twin(twin, case, td, extra=(), files=None):
    tfiles, cfiles = files if files else util.write_case_sams(case, td)
    out, logf, pile = (os.path.join(td, x) for x in ("o.np", "o.f", "o.k"))
    cmd = [twin, "-t", ",".join(tfiles), "-o", out, "-f", logf, "-k", pile, "-v"] + case.ref_args() + list(extra)
    if any(c != "null" for c in cfiles):
        cmd += ["-c", ",".join(cfiles)]
    if case.bed:
        bedf = os.path.join(td, "x.bed")
        util.write_case_bed(case, bedf)
        cmd += ["-E", bedf]
    r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    return out, logf, pile, r.stderr


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_host_program_matches_reference_files(twin, case, tmp_path):
    out, logf, pile, err = run_twin(twin, case, str(tmp_path))
    meta, gold = util.golden(case)
    assert open(out).read().split("\n")[:-1] == gold                     # narrowPeak, every column
    assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])         # the whole -f file
    assert _sha(pile, True) == (meta["pile_sha256"], meta["pile_lines"])  # the whole -k file (minus the path lines)
    lam = [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)]
    fac = [float(x) for x in re.findall(r"Scaling factor for control pileup: ([0-9.]+)", err)]
    assert lam == meta["lambda"] and fac == meta["factor"]
    assert int(re.search(r"Genome length: (\d+)bp", err).group(1)) == meta["genome_len"]
    assert int(re.search(r"Peaks identified: (\d+)", err).group(1)) == meta["peaks"]
    assert int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)) == meta["peak_bp"]
    assert ("All q-values are 1" in err) == meta["all_q_one"]
    assert len(re.findall(r"prevented from extending", err)) == meta["clamp_warnings"]


# ---- host-side options on SAM files with unpaired / discordant alignments, PCR duplicates,
# ---- quality strings and low MAPQ (tests/hostcases.py): -y -w -x -m -e -X -r -R
from hostcases import HOST_CASES, write_host_sams, host_cmd  # noqa: E402
import json  # noqa: E402


def host_golden(h):
    with open(os.path.join(util.GOLDEN, h.name + ".json")) as f:
        meta = json.load(f)
    p = os.path.join(util.GOLDEN, h.name + ".narrowPeak")
    peaks = open(p).read() if os.path.exists(p) else None
    return meta, peaks


def check_host_case(binary, h, td, exact=True):
    tfiles, cfiles = write_host_sams(h, td)
    bedf = None if "--threads" in h.args else os.path.join(td, "o.bed")     # -b keeps the decode on one thread
    cmd, out, logf, dupf = host_cmd(binary, h, td, tfiles, cfiles, bed=bedf)
    r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    meta, peaks = host_golden(h)
    if bedf:
        assert _sha(bedf) == (meta["bed_sha256"], meta["bed_lines"])               # the whole -b file
    err = r.stderr.replace(td, "@")
    if h.dups_log:
        assert _sha(dupf, True) == (meta["dups_sha256"], meta["dups_lines"])      # the whole -R log
    if exact:
        assert err == meta["stderr"]                                   # the complete -v text
        if peaks is not None:
            assert open(out).read() == peaks
        assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])
    return out, logf, err, meta, peaks


@pytest.mark.parametrize("h", HOST_CASES, ids=lambda h: h.name)
def test_host_options_match_reference(twin, h, tmp_path):
    check_host_case(twin, h, str(tmp_path))


def test_host_bam_and_gz_inputs(twin, tmp_path):
    """BAM (BGZF, raw quality bytes) and gzip-compressed SAM give what plain SAM gives -- with -r,
    where the order of evaluation depends on the quality sums read from either format."""
    h = [x for x in HOST_CASES if x.name == "host_r_y"][0]
    td = str(tmp_path)
    tfiles, cfiles = write_host_sams(h, td)
    bam = os.path.join(td, "t.bam")
    util.sam_to_bam(tfiles[0], bam)
    gz = os.path.join(td, "t.sam.gz")
    import gzip
    with open(tfiles[0], "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    meta, peaks = host_golden(h)
    # (file, host threads, environment): BAM through zlib's gzread, then through the threaded BGZF
    # inflater with batches so small that records straddle many of them
    mt = dict(os.environ, GB_THREAD_MIN_BYTES="1", GB_BGZF_BATCH_BYTES="150000")
    for path, threads, env in ((bam, 1, None), (bam, 3, mt), (bam, 8, dict(mt, GB_BGZF_BATCH_BYTES="1")), (gz, 4, mt)):
        cmd, out, logf, dupf = host_cmd(twin, h, td, [path], cfiles)
        cmd += ["--threads", str(threads)]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert open(out).read() == peaks
        assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])
        assert _sha(dupf, True) == (meta["dups_sha256"], meta["dups_lines"])
        want = meta["stderr"].replace("@/mt0.sam", path.replace(td, "@"))
        if path == bam:
            want = want.replace("SAM records analyzed", "BAM records analyzed")
        assert r.stderr.replace(td, "@") == want


def test_host_errors(twin, tmp_path):
    td = str(tmp_path)
    r = subprocess.run([twin, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error! Need input/output files" in r.stderr
    bad = os.path.join(td, "bad.sam")
    open(bad, "w").write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:1000\n")
    r = subprocess.run([twin, "-t", bad, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "not sorted by queryname" in r.stderr
    emp = os.path.join(td, "empty.sam")
    open(emp, "w").write("@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:chr1\tLN:1000\n")
    r = subprocess.run([twin, "-t", emp, "-o", os.path.join(td, "x")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Experimental sample has no analyzable fragments" in r.stderr
    sat = os.path.join(td, "sat.sam")                    # 32768 identical fragments: one more than an int16 counter holds
    with open(sat, "w") as f:
        f.write("@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:chr1\tLN:5000\n")
        for i in range(32768):
            f.write("r%d\t99\tchr1\t1001\t42\t50M\t=\t1201\t250\t*\t*\n" % i)
            f.write("r%d\t147\tchr1\t1201\t42\t50M\t=\t1001\t-250\t*\t*\n" % i)
    # the reference drops the last one and says so under -v (saveInterval 2558-2573); so does the host program
    r = subprocess.run([twin, "-t", sat, "-o", os.path.join(td, "x"), "-v"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    warn = [l for l in r.stderr.split("\n") if "skipped due to" in l]
    assert len(warn) == 1 and "Read r32767, alignment at (chr1, 1000-1250) skipped due to overflow" in warn[0]


@pytest.mark.parametrize("threads", [2, 5])
def test_threaded_decode_same_output(twin, threads, tmp_path, monkeypatch):
    """--threads N: a plain SAM file cut into N pieces at read-name boundaries and decoded by N
    threads gives the files of the sequential decode -- i.e. the reference's, byte for byte --
    including the -r duplicate log and the -v text (warnings replayed in file order)."""
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    names = ["c2_ctrl_q", "c3_atac_q", "c5_multimap_ctrl_p", "c4_fisher_q", "bed_fisher_q"]
    for name in names:
        case = BY_NAME[name]
        td = str(tmp_path / (name + str(threads)))
        os.makedirs(td)
        out, logf, pile, err = run_twin(twin, case, td, extra=["--threads", str(threads)])
        meta, gold = util.golden(case)
        assert open(out).read().split("\n")[:-1] == gold
        assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])
        assert len(re.findall(r"prevented from extending", err)) == meta["clamp_warnings"]
    for h in HOST_CASES:
        if h.name in ("host_y", "host_x", "host_r", "host_r_y", "host_r_x", "host_r_multimap", "host_m_e"):
            td = str(tmp_path / (h.name + str(threads)))
            os.makedirs(td)
            h2 = type(h)(h.name, h.case, list(h.args) + ["--threads", str(threads)], h.dups_log)
            check_host_case(twin, h2, td)


def test_host_stdin_gzip_out_and_bed(twin, tmp_path):
    """-t - (SAM piped in; spooled, since every input is read twice), -z (gzip-compressed outputs)
    and -b (the interval file) against the reference's files."""
    import gzip
    case = BY_NAME["c2_ctrl_p"]
    td = str(tmp_path)
    tfiles, cfiles = util.write_case_sams(case, td)
    meta, gold = util.golden(case)
    out, bed = os.path.join(td, "o.np"), os.path.join(td, "o.bed")
    with open(tfiles[0], "rb") as f:
        r = subprocess.run([twin, "-t", "-", "-c", cfiles[0], "-o", out, "-b", bed, "-z", "-v"] + case.ref_args(),
                           stdin=f, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert "Processing experimental file #0: -\n" in r.stderr
    assert gzip.open(out + ".gz", "rt").read().split("\n")[:-1] == gold
    # -b: one line per interval, in file order, "<read>_<count>_<E|C>_<sample>"
    want = []
    for arr, tag in zip(util.case_inputs(case)[0][:2], "EC"):
        want += ["chr%d\t%d\t%d\t%d_%s_0" % (a + 1, b, c2, k, tag) for a, b, c2, k in arr.tolist()]
    got = []
    for l in gzip.open(bed + ".gz", "rt").read().split("\n")[:-1]:
        f = l.split("\t")
        nm = f[3].split("_")
        got.append("%s\t%s\t%s\t%s_%s_%s" % (f[0], f[1], f[2], nm[-3], nm[-2], nm[-1]))
    assert got == want
    with gzip.open(os.path.join(td, "t.gz"), "wb") as g, open(tfiles[0], "rb") as f:
        g.write(f.read())
    with open(os.path.join(td, "t.gz"), "rb") as f:
        r = subprocess.run([twin, "-t", "-", "-o", out], stdin=f, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Cannot pipe in gzip-compressed file" in r.stderr


@pytest.mark.parametrize("gpus", [2, 3])
def test_sharded_contexts_same_output(twin, gpus, tmp_path, monkeypatch):
    """--gpus N: the host program's own dispatcher -- chromosomes sharded over N engine contexts, the
    per-chromosome sums added on the host, one p-value histogram handed to every context, peaks
    merged in chromosome order -- reproduces the reference's files for every seeded case (narrowPeak,
    -f, -k byte for byte; lambda, scale factor, genome length, peak counts in the -v text), also
    together with --threads.  (Here the contexts are oracle contexts; on the GPU box they are devices.)"""
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    for case in CASES:
        td = str(tmp_path / (case.name + str(gpus)))
        os.makedirs(td)
        extra = ["--gpus", str(gpus)] + (["--threads", "3"] if case.name in ("c2_ctrl_q", "c4_fisher_q") else [])
        out, logf, pile, err = run_twin(twin, case, td, extra=extra)
        meta, gold = util.golden(case)
        assert open(out).read().split("\n")[:-1] == gold, case.name
        assert _sha(logf) == (meta["log_sha256"], meta["log_lines"]), case.name
        assert _sha(pile, True) == (meta["pile_sha256"], meta["pile_lines"]), case.name
        lam = [float(x) for x in re.findall(r"Background pileup value: ([0-9.]+)", err)]
        fac = [float(x) for x in re.findall(r"Scaling factor for control pileup: ([0-9.]+)", err)]
        assert lam == meta["lambda"] and fac == meta["factor"], case.name
        assert int(re.search(r"Genome length: (\d+)bp", err).group(1)) == meta["genome_len"]
        assert int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)) == meta["peak_bp"]
        assert ("All q-values are 1" in err) == meta["all_q_one"]
    for h in HOST_CASES:
        if h.name in ("host_r_y", "host_x", "host_m_e", "host_X"):
            td = str(tmp_path / (h.name + str(gpus)))
            os.makedirs(td)
            h2 = type(h)(h.name, h.case, list(h.args) + ["--gpus", str(gpus)], h.dups_log)
            check_host_case(twin, h2, td)


PONLY = sorted(f[:-5] for f in os.listdir(util.GOLDEN) if f.startswith("ponly_") and f.endswith(".json"))


def test_peaks_only(twin, tmp_path):
    """-P (peaks from a -f log: findPeaksOnly 5243, callPeaksLog 1277): the host program parses its own
    -f log -- byte for byte the reference's -- and calls peaks on the log's -log(p) / -log(q) columns
    with new thresholds, -e, and new -E regions; narrowPeak files, genome length, peak counts and
    warnings equal what the unmodified reference produced with the same arguments
    (tests/golden/ponly_*, make_golden_peaksonly.py).  Also gzip input and the error texts."""
    import gzip
    import json
    logs = {}
    for name in PONLY:
        meta = json.load(open(os.path.join(util.GOLDEN, name + ".json")))
        cname = meta["case"]
        if cname not in logs:
            td = str(tmp_path / cname)
            os.makedirs(td)
            logs[cname] = run_twin(twin, BY_NAME[cname], td)[1]
            cm, _ = util.golden(BY_NAME[cname])
            assert _sha(logs[cname]) == (cm["log_sha256"], cm["log_lines"])
        out = str(tmp_path / (name + ".np"))
        cmd = [twin, "-P", "-f", logs[cname], "-o", out, "-v"] + meta["args"]
        if meta["bed_case"]:
            bedf = str(tmp_path / (name + ".bed"))
            util.write_case_bed(BY_NAME[meta["bed_case"]], bedf)
            cmd += ["-E", bedf]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, (name, r.stderr)
        gold = open(os.path.join(util.GOLDEN, name + ".narrowPeak")).read()
        assert open(out).read() == gold, name
        err = r.stderr
        assert int(re.search(r"Genome length: (\d+)bp", err).group(1)) == meta["genome_len"], name
        assert int(re.search(r"Peaks identified: (\d+)", err).group(1)) == meta["peaks"], name
        assert int(re.search(r"Peaks identified: \d+ \((\d+)bp\)", err).group(1)) == meta["peak_bp"], name
        assert ("Skipping given BED regions" in err) == meta["warn_bed"], name
        assert len(re.findall(r"Skipping chromosome", err)) == meta["warn_chr"], name
    # gzip-compressed log in, gzip-compressed peaks out
    lg = logs["c2_ctrl_q"]
    gz = str(tmp_path / "log.gz")
    with open(lg, "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    out = str(tmp_path / "z.np")
    r = subprocess.run([twin, "-P", "-f", gz, "-o", out, "-q", "0.05", "-z"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert gzip.open(out + ".gz", "rt").read() == open(os.path.join(util.GOLDEN, "ponly_c2_same.narrowPeak")).read()   # openWrite 5088: ".gz" appended
    # a -p log has no -log(q) column; a log without header fields; missing -f
    r = subprocess.run([twin, "-P", "-f", logs["c1_smoke"], "-o", out, "-q", "0.05"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error! -log(q): cannot find field in header of bedgraph-ish log file" in r.stderr
    bad = str(tmp_path / "bad.f")
    open(bad, "w").write("chr\tstart\tend\nchr1\t0\t10\n")
    r = subprocess.run([twin, "-P", "-f", bad, "-o", out], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error! -log(p): cannot find field in header" in r.stderr
    r = subprocess.run([twin, "-P", "-o", out], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Need input/output files" in r.stderr


@pytest.mark.parametrize("threads", [1, 3])
def test_host_saturation_rule(twin, threads, tmp_path, monkeypatch):
    """saveInterval 2558-2573 through the host program: 48 466 alignments of tests/satcase.py are dropped by the
    reference once its int16 counters saturate, in file order.  narrowPeak, -f and -k equal the reference's, and
    under -v the WHOLE stderr text does, the 48 466 "skipped due to overflow / underflow" lines with their read
    names included -- also when three decode workers share the file (their records reach the engine in file order)."""
    import satcase
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    td = str(tmp_path)
    sam = os.path.join(td, "t.sam")
    satcase.write_sam(sam)
    meta = json.load(open(os.path.join(util.GOLDEN, "sat_hot.json")))
    out, logf, pile = (os.path.join(td, x) for x in ("o.np", "o.f", "o.k"))
    r = subprocess.run([twin, "-t", sam, "-o", out, "-f", logf, "-k", pile, "-v", "--threads", str(threads)] + satcase.ARGS,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out).read() == open(os.path.join(util.GOLDEN, "sat_hot.narrowPeak")).read()
    assert _sha(logf) == (meta["log_sha256"], meta["log_lines"])
    assert _sha(pile, True) == (meta["pile_sha256"], meta["pile_lines"])
    assert hashlib.sha256(r.stderr.replace(sam, "T.sam").encode()).hexdigest() == meta["verbose_sha256"]


def test_threaded_decode_loosely_grouped(twin, tmp_path, monkeypatch):
    """A file whose alignment sets are interleaved with unmapped lines of OTHER names (A, x unmapped, A): readSAM
    never looks at the name of a dropped line, so both A lines are one set.  The threaded decoder cuts the file
    between kept-record sets -- same output with 1 and with 5 workers (and as the reference, where it is built)."""
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    case = BY_NAME["c2_ctrl_p"]
    td = str(tmp_path)
    tfiles, cfiles = util.write_case_sams(case, td)
    loose = os.path.join(td, "loose.sam")
    with open(tfiles[0]) as f, open(loose, "w") as g:
        n = 0
        for line in f:
            g.write(line)
            if not line.startswith("@"):
                n += 1
                if n % 2 == 1:                 # between the two mates of every pair
                    g.write("u%d\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n" % n)
    outs = []
    for th in (1, 5):
        o = os.path.join(td, "o%d.np" % th)
        subprocess.check_call([twin, "-t", loose, "-c", cfiles[0], "-o", o, "-S", "--threads", str(th)] + case.ref_args())
        outs.append(open(o).read())
    assert outs[0] == outs[1] and len(outs[0]) > 0
    ref = os.path.join(util.ORACLE_DIR, "_ref", "Genrich")
    if os.path.exists(ref):
        o = os.path.join(td, "ref.np")
        subprocess.check_call([ref, "-t", loose, "-c", cfiles[0], "-o", o, "-S"] + case.ref_args(), stderr=subprocess.DEVNULL)
        assert open(o).read() == outs[0]


REF_BIN = os.path.join(util.ORACLE_DIR, "_ref", "Genrich")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/Genrich not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(5000, 5012))
def test_random_cases_against_the_reference_binary(twin, seed, tmp_path, monkeypatch):
    """Seeded random cases (tests/fuzzcases.py) with random host options on mutated SAM files (singletons, discordant
    pairs, duplicates, low MAPQ): the unmodified reference binary and the host program run on the same files, here,
    and every output is compared byte for byte -- narrowPeak, -f, -k, -b, -R and the whole -v text.  Odd seeds decode
    with three threads."""
    import hostcases
    from fuzzcases import random_case, random_host_options
    monkeypatch.setenv("GB_THREAD_MIN_BYTES", "1")
    case = random_case(seed)
    extra = random_host_options(seed, case)
    td = str(tmp_path)
    tf, cf = util.write_case_sams(case, td)

    def mutated(p, k):
        q = p.replace(".sam", ".m.sam")
        hostcases.mutate_sam(p, q, seed + k)
        return q
    tf = [mutated(p, i) for i, p in enumerate(tf)]
    cf = [c if c == "null" else mutated(c, 100 + i) for i, c in enumerate(cf)]
    res = []
    for exe, tag in ((REF_BIN, "A"), (twin, "B")):
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
        if tag == "B" and seed % 2:
            cmd += ["--threads", "3"]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
        res.append([r.returncode] + [open(p, "rb").read() if os.path.exists(p) else None for p in f.values()] +
                   [r.stderr.replace(tag + ".", "X.")])
    for name, a, b in zip(("exit code", "narrowPeak", "-f", "-k", "-R", "-b", "-v text"), res[0], res[1]):
        assert a == b, (name, case, extra)
