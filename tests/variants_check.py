"""Child process of tests/test_gpu_zz_variants.py: one non-default kernel variant (environment
knobs) against the plain scatter + streaming scan on the seeded cases, the edge inputs and a
200 Mbp sample -- same bits everywhere (interval ends, pileup floats, lambda, scale factor, peaks).

    python tests/variants_check.py GR_FUSED_RANK=1 [GR_FB_SLOTS=1 ...]

Runs in a process of its own so that a kernel that faults or hangs on the device takes only
itself down (the parent gives it a time limit)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import util                                    # noqa: E402
from cases import CASES                        # noqa: E402
from genrich_b200 import capi, host            # noqa: E402
from genrich_b200.synth import Workload        # noqa: E402

PLAIN = {"GR_FUSED": "0", "GR_SB_MIN": "1000000000"}
FUSED = {"GR_FUSED": "1", "GR_FUSED_MIN": "1"}


def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def run(api, env, chrom_len, par, inputs, chunk=30011):
    for k, v in env.items():
        os.environ[k] = v
    try:
        ctx = capi.Context(api, chrom_len, par)
        res = host.run_replicates(ctx, inputs, chunk=chunk)
        return res, [[ctx.fetch(w, 0, c) for c in range(len(chrom_len))] for w in (0, 1, 2)]
    finally:
        for k in env:
            os.environ.pop(k, None)


def same(a, b, what):
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


def main():
    variant = dict(FUSED)
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        variant[k] = v
    if os.environ.get("GR_EMU_AS_CUDA"):       # development aid: the CPU-emulated build of the library (tests/emu)
        capi._cuda_api = capi.Api(os.environ.get("GR_EMU_LIB") or os.path.join(ROOT, "tests", "emu", "_build", "libgenrich_emu.so"), "gr_")
    api = capi.load_cuda()
    n = 0
    only = set(filter(None, os.environ.get("GR_VARIANT_CASES", "").split(",")))      # the CPU suite runs a subset
    for case in CASES:
        if case.bed or (only and case.name not in only):
            continue
        inputs = [list(r) for r in util.case_inputs(case)]
        extra = np.array([[0, 1000, 41000, 2], [0, 8000, 3 * 8192 + 5, 1], [0, 8191, 8193, 4], [0, 8192, 8192, 3],
                          [0, 16383, 16384, 5], [0, 0, 8192, 6]], np.int32)
        extra = extra[extra[:, 2] <= case.chrom_len[0]]
        inputs[0][0] = np.concatenate([inputs[0][0], extra])
        par = util.case_params(case)
        same(run(api, PLAIN, case.chrom_len, par, inputs), run(api, variant, case.chrom_len, par, inputs), case.name)
        n += 1
    L = [5000, 8192, 8191, 1, 20000, 16384, 16383]
    recs = np.array([
        [0, 0, 5000, 1], [0, -50, 10, 2], [0, 4990, 6000, 3], [1, 0, 1, 1], [1, 8191, 8192, 1],
        [2, 8190, 8191, 10], [2, 0, 8191, 8], [3, 0, 1, 1], [4, 100, 100, 5],
        [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6], [4, 300, 900, 6],
        [4, 19999, 25000, 4], [5, 0, 16384, 1], [5, 8191, 8192, 2], [5, 8192, 8193, 2], [5, 16383, 16384, 3],
        [6, 0, 16383, 1], [6, 8100, 16383, 2], [6, 16382, 16383, 3],
    ], dtype=np.int32)
    par = capi.make_params(p=0.2, min_auc=0.5, keep_pileups=True)
    a, b = run(api, PLAIN, L, par, [(recs, None)]), run(api, variant, L, par, [(recs, None)])
    same(a, b, "edge")
    assert b[0].sample_stats[0].n_clamped == 3
    # 200 Mbp / 4 M + 4 M fragments with hot spots: thousands of events in one block, many pages per owner
    # (under emulation: a tenth of it)
    scale = int(os.environ.get("GR_EMU_SCALE", "10")) if os.environ.get("GR_EMU_AS_CUDA") else 1
    L = [x // scale for x in (60_000_000, 50_000_000, 40_000_000, 30_000_000, 20_000_000)]
    t = Workload(L, 4_000_000 // scale, 101, enrich=0.5, spacing=400000 // scale, sigma=60.0).fragments()
    c = Workload(L, 4_000_000 // scale, 102, enrich=0.0).fragments()
    par = capi.make_params(p=0.01, min_auc=20.0)
    a = run(api, {"GR_FUSED": "1"}, L, par, [(t, c)], chunk=1 << 22)       # the default path, validated against the dense one
    b = run(api, variant, L, par, [(t, c)], chunk=1 << 22)
    same(a, b, "large")
    assert b[0].sample_stats[0].frag_len == float(np.sum((t[:, 2] - t[:, 1]).astype(np.int64)))
    assert len(b[0].peaks) > 100 // scale
    print("variant %s: %d seeded cases, edge inputs and the 200 Mbp sample identical to the default path" %
          (" ".join(sys.argv[1:]), n))


if __name__ == "__main__":
    main()
