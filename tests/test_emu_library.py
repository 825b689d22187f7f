"""The whole CUDA library on the CPU.  tests/emu compiles libgenrich_cuda's own sources -- every kernel
AND the host logic behind the C-ABI (gr_api.cu: buffer sizes, launch arguments, stage order, retries)
-- against a lock-step emulation of warps / CTAs and a synchronous stand-in for the CUDA runtime
(tests/emu/fake/cuda_runtime.h; "device" buffers are filled with 0xCD, not zeros), into
tests/emu/_build/libgenrich_emu.so.  It is driven through the same ctypes binding as the CUDA library
and compared with the pinned oracle: peaks, interval partitions, pileup floats bit for bit, and
-log10 p / q (same glibc on both sides here, so these are bit-exact too).

TEST INFRASTRUCTURE: nothing in the product loads this library (the product path needs a GPU and
fails without one, tests/test_abi.py).  What it buys: kernels and host plumbing written where no GPU
is at hand are exercised end to end before they ever reach a device; what it cannot see: stream ordering, memory-model races, performance."""
import os
import subprocess

import numpy as np
import pytest

import util
from cases import BY_NAME
from genrich_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
LIB = os.path.join(EMU, "_build", "libgenrich_emu.so")

FUSED = {"GR_FUSED": "1", "GR_FUSED_MIN": "1"}
MODES = {
    "default_small": {},                                     # plain scatter + streaming scan (small samples)
    "default_fused": FUSED,                                  # the path the bench takes: the sample picks the scan form
    "fused_cta": dict(FUSED, GR_FUSED_CTA="1"),              # k_fb_scan without -E regions
    "fused_rank": dict(FUSED, GR_FUSED_RANK="1"),            # k_fr_scan whatever the blocks hold
}


@pytest.fixture(scope="module")
def emu_api():
    subprocess.check_call(["make", "-s", "-C", EMU, "_build/libgenrich_emu.so"])
    return capi.Api(LIB, "gr_")


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def _compare(case, api, env, monkeypatch, packed=False):
    inputs = util.case_inputs(case)
    ctx_o, res_o, _ = util.run_case(util.oracle_api(), case, inputs=inputs)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if packed:
        from genrich_b200 import host
        ctx_g = capi.Context(api, case.chrom_len, util.case_params(case))
        if case.bed:
            ctx_g.set_exclusions(case.bed)
        res_g = host.run_replicates(ctx_g, inputs, chunk=20011, packed=packed)
    else:
        ctx_g, res_g, _ = util.run_case(api, case, inputs=inputs)
    a, b = res_g.peaks, res_o.peaks
    assert len(a) == len(b)
    for f in ("chrom", "start", "end", "summit"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("pval", "qval", "auc"):
        assert np.max(np.abs(a[f].astype(np.float64) - b[f])) <= 1e-4 if len(a) else True, f
    for sa, sb in zip(res_g.sample_stats, res_o.sample_stats):
        assert (sa.frag_len, sa.ctrl_frag, sa.n_expt, sa.n_ctrl, sa.n_pval, sa.n_clamped) == \
               (sb.frag_len, sb.ctrl_frag, sb.n_expt, sb.n_ctrl, sb.n_pval, sb.n_clamped)
        assert _bits(sa.lambda_) == _bits(sb.lambda_) and _bits(sa.factor) == _bits(sb.factor)
    nrep = len(case.reps)
    for which, rep in ((0, 0), (1, 0), (2, nrep - 1)):         # last sample's pileups, last replicate's p intervals
        for c in range(len(case.chrom_len)):
            x, y = ctx_g.fetch(which, rep, c), ctx_o.fetch(which, rep, c)
            assert (x is None) == (y is None), (which, c)
            if x is not None:
                assert np.array_equal(x.end, y.end), (which, c)
                if which < 2:
                    assert np.array_equal(_bits(x.val), _bits(y.val)), (which, c)
                else:
                    fin = np.isfinite(y.val) & (y.val < 3e38)
                    assert np.max(np.abs(x.val[fin].astype(np.float64) - y.val[fin]), initial=0.0) <= 1e-4, (which, c)
    assert ctx_g.kernel_launches() > 0
    return ctx_g.kernel_launches()


@pytest.mark.parametrize("mode", sorted(MODES))
def test_emulated_library_matches_oracle(emu_api, mode, monkeypatch):
    """Treatment + control, -q (every stage incl. BH), through each formulation of the per-base pass."""
    _compare(BY_NAME["c2_ctrl_q"], emu_api, MODES[mode], monkeypatch)


@pytest.mark.parametrize("name,mode,packed", [
    ("c4_fisher_q", "default_fused", False),     # three replicates, Fisher combine
    ("c5_multimap_ctrl_p", "default_fused", True),   # fractional weights, 8-byte packed records
    ("c3_atac_q", "default_fused", 6),           # ATAC intervals, 6-byte packed records
    ("bed_ctrl_q", "default_fused", False),      # -E regions: marks exist in k_fb_scan only
])
def test_emulated_library_other_shapes(emu_api, name, mode, packed, monkeypatch):
    _compare(BY_NAME[name], emu_api, MODES[mode], monkeypatch, packed=packed)


def test_saturation_rule_under_emulation(emu_api, monkeypatch):
    """k_sat_resolve (the reference's int16 saturation skips, saveInterval 2558-2573) through the C-ABI: the same
    records dropped as by the oracle -- which is pinned to the unmodified reference on this very case
    (tests/test_oracle_pin.py) -- in one push and in many, 16-byte and packed records."""
    assert util.check_saturation(emu_api) > 0
    assert util.check_saturation(emu_api, packed=True, chunk=50021) > 0


def test_host_program_over_emulated_devices(emu_api, tmp_path):
    """The host program linked against the emulated CUDA library, --gpus 3 over three emulated devices
    (every device allocation remembers its device; a copy or memset issued while another device is
    current aborts): chromosomes sharded over three contexts of the CUDA library -- per-chromosome sums
    added on the host, one p-value histogram through gr_bh_*_host, peaks merged in chromosome order --
    reproduce the reference's narrowPeak, -f and -k files byte for byte."""
    import hashlib
    subprocess.check_call(["make", "-s", "-C", EMU, "_build/genrich-b200-emu"])
    cli = os.path.join(EMU, "_build", "genrich-b200-emu")
    for name in ("c2_ctrl_q", "c4_fisher_q", "bed_ctrl_q"):
        case = BY_NAME[name]
        td = str(tmp_path / name)
        os.makedirs(td)
        tfiles, cfiles = util.write_case_sams(case, td)
        out, logf, pile = (os.path.join(td, x) for x in ("o.np", "o.f", "o.k"))
        cmd = [cli, "-t", ",".join(tfiles), "-o", out, "-f", logf, "-k", pile, "--gpus", "3"] + case.ref_args()
        if any(c != "null" for c in cfiles):
            cmd += ["-c", ",".join(cfiles)]
        if case.bed:
            bedf = os.path.join(td, "x.bed")
            util.write_case_bed(case, bedf)
            cmd += ["-E", bedf]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, env=dict(os.environ, EMU_DEVICES="4"))
        assert r.returncode == 0, r.stderr
        meta, gold = util.golden(case)
        assert open(out).read().split("\n")[:-1] == gold, name
        h = hashlib.sha256()
        nl = 0
        for line in open(logf, "rb"):
            h.update(line)
            nl += 1
        assert (h.hexdigest(), nl) == (meta["log_sha256"], meta["log_lines"]), name


def test_peaks_only_over_emulated_library(emu_api, tmp_path):
    """-P with the host program over the emulated CUDA library: gr_load_pvalues + the K8 kernels on the
    -log(p) / -log(q) columns of a -f log, against the unmodified reference's -P files (tests/golden/ponly_*)."""
    import json
    subprocess.check_call(["make", "-s", "-C", EMU, "_build/genrich-b200-emu"])
    cli = os.path.join(EMU, "_build", "genrich-b200-emu")
    case = BY_NAME["c2_ctrl_q"]
    td = str(tmp_path)
    tfiles, cfiles = util.write_case_sams(case, td)
    logf = os.path.join(td, "o.f")
    subprocess.run([cli, "-t", ",".join(tfiles), "-c", ",".join(cfiles), "-o", os.path.join(td, "o.np"), "-f", logf] + case.ref_args(),
                   check=True, stderr=subprocess.DEVNULL)
    for name in ("ponly_c2_same", "ponly_c2_p_strict", "ponly_c2_q_skipchr", "ponly_c2_newbed", "ponly_c2_newbed_q"):
        meta = json.load(open(os.path.join(util.GOLDEN, name + ".json")))
        out = os.path.join(td, name + ".np")
        cmd = [cli, "-P", "-f", logf, "-o", out] + meta["args"]
        if meta["bed_case"]:
            bedf = os.path.join(td, name + ".bed")
            util.write_case_bed(BY_NAME[meta["bed_case"]], bedf)
            cmd += ["-E", bedf]
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, (name, r.stderr)
        assert open(out).read() == open(os.path.join(util.GOLDEN, name + ".narrowPeak")).read(), name


def test_bench_e2e_call_pattern(emu_api, monkeypatch):
    """The call sequence of bench.py's e2e arm, which no other test drives: 6-byte records in PINNED host
    buffers, both samples of the NEXT step sent ahead (gr_prefetch_packed6, two slots) while the current
    step runs without host round trips (sample_pileup_async, replicate_finish_device), gr_reset between
    steps -- three steps give the peaks of the plain synchronous path, and so does the 8-byte variant."""
    import ctypes as C
    from genrich_b200 import host
    for k, v in FUSED.items():
        monkeypatch.setenv(k, v)
    case = BY_NAME["c2_ctrl_p"]
    (t, c), = [r[:2] for r in util.case_inputs(case)]
    par = util.case_params(case)
    ref = host.run_replicates(capi.Context(emu_api, case.chrom_len, par), [(t, c)]).peaks
    assert len(ref) > 0
    ctx = capi.Context(emu_api, case.chrom_len, par)
    lib = C.CDLL(LIB)
    lib.gr_pinned_alloc.restype = C.c_void_p
    lib.gr_pinned_alloc.argtypes = [C.c_size_t]
    lib.gr_pinned_free.argtypes = [C.c_void_p]
    layout = ctx.pack6_layout()
    assert layout is not None
    bufs = []

    def pinned(a):
        p = lib.gr_pinned_alloc(a.nbytes)
        C.memmove(p, a.ctypes.data, a.nbytes)
        bufs.append(p)
        return p
    t6, rest = host.pack6_records(t, layout, case.chrom_len)
    c6, rest2 = host.pack6_records(c, layout, case.chrom_len)
    assert not len(rest) and not len(rest2)
    t8, _ = host.pack_records(t)
    c8, _ = host.pack_records(c)
    pt6, pc6, pt8, pc8 = pinned(t6), pinned(c6), pinned(t8), pinned(c8)
    glen = int(sum(case.chrom_len))
    for six in (True, False):
        pt, pc = (pt6, pc6) if six else (pt8, pc8)
        push = ctx.push_packed6_ptr if six else ctx.push_packed_ptr
        pre = ctx.prefetch_packed6_ptr if six else ctx.prefetch_packed_ptr
        for step in range(3):
            ctx.reset()
            ctx.sample_begin(False, None)
            push(pt, len(t))
            ctx.sample_pileup_async()
            ctx.sample_begin(True)
            push(pc, len(c))
            ctx.sample_pileup_async()
            ctx.replicate_finish_device(True, glen)
            pre(pt, len(t))                                  # the next step's samples travel under this step's kernels
            pre(pc, len(c))
            peaks, _ = ctx.call_peaks()
            assert peaks.tobytes() == ref.tobytes(), (six, step)
    for p in bufs:
        lib.gr_pinned_free(p)


@pytest.mark.parametrize("seed", range(24))
def test_emulated_library_random_cases(emu_api, seed, monkeypatch):
    """Seeded random cases through the whole library (every kernel, the host logic of gr_api.cu) against the oracle:
    interval ends, pileup floats, lambda, scale factor, peak coordinates bit for bit; -log10 p / q within 1e-4.
    The per-base pass alternates between the plain scatter path, the fused scan with the form chosen on the device,
    and each form forced; every third case travels as packed records."""
    from fuzzcases import random_case
    case = random_case(seed, holes_with_multimap=False)
    mode = sorted(MODES)[seed % len(MODES)]
    if case.bed and mode == "default_small":
        mode = "default_fused"                       # -E regions exist in the fused scan only (the library routes them there anyway)
    try:
        util.run_case(util.oracle_api(), case)
    except capi.GenrichError as want:                # e.g. a replicate left without fragments: the same refusal is expected
        with pytest.raises(capi.GenrichError) as got:
            _compare(case, emu_api, MODES[mode], monkeypatch, packed=(seed % 3 == 0))
        assert got.value.status == want.status
        return
    _compare(case, emu_api, MODES[mode], monkeypatch, packed=(seed % 3 == 0))


REF_BIN = os.path.join(util.ORACLE_DIR, "_ref", "Genrich")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/Genrich not built (needs /root/reference)")
@pytest.mark.parametrize("seed,gpus", [(6000, 0), (6002, 0), (6007, 0), (6013, 0), (6020, 0), (6044, 0), (6101, 3), (6105, 3)])
def test_emulated_cli_random_cases_against_the_reference_binary(emu_api, seed, gpus, tmp_path):
    """The host program over the CUDA library (compiled for the CPU) and the unmodified reference binary on the same
    mutated SAM files of a seeded random case with random host options: narrowPeak, -f, -k, -b, -R and the whole -v
    text byte for byte; two of the cases sharded over three emulated devices.  (Offline this ran over 90 seeds.)"""
    import hostcases
    from fuzzcases import random_case, random_host_options
    subprocess.check_call(["make", "-s", "-C", EMU, "_build/genrich-b200-emu"])
    cli = os.path.join(EMU, "_build", "genrich-b200-emu")
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
    for exe, tag in ((REF_BIN, "A"), (cli, "B")):
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
        r = subprocess.run(cmd, stderr=subprocess.PIPE, text=True, env=env, timeout=300)
        res.append([r.returncode] + [open(p, "rb").read() if os.path.exists(p) else None for p in f.values()] +
                   [r.stderr.replace(tag + ".", "X.")])
    for name, a, b in zip(("exit code", "narrowPeak", "-f", "-k", "-R", "-b", "-v text"), res[0], res[1]):
        assert a == b, (name, case, extra)
