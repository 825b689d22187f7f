"""BASELINE.json configs[1] at its FULL size (hg38-sized genome, 25 chromosomes, 50 M treatment +
50 M control fragments) through size-independent properties -- the oracle would need minutes
for this, the properties need none:

  * sum(len * val) of the two pileups == total fragment bp (exact: integer weights), hence
    lambda and the scale factor are the correctly rounded quotients;
  * the p-value intervals of every chromosome tile it: ends strictly increase, the last end is
    the chromosome length, so sum(interval lengths) == genome length;
  * every experimental / control break is a p-interval break (union, savePval 1768-1791), and the
    union holds no other break;
  * peaks lie inside their chromosomes, are ordered, summits inside, -log10 p above the threshold;
  * the fused formulation (delta cells in shared memory only) and the dense one (delta array in HBM)
    give byte-identical peak records and p-interval ends; packed 8-byte records through the
    no-round-trip path give the same peaks as int32 x 4 records through the synchronous one;
  * a second run is bit-identical (integer atomics, fixed-point sums: nothing depends on order).
"""
import os

import numpy as np
import pytest

from genrich_b200 import capi, host
from genrich_b200.synth import Workload

pytestmark = pytest.mark.gpu

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
        138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
        83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
N = 50_000_000


def _fragments(n, seed, enrich):
    import threading
    w = Workload(HG38, n, seed, enrich=enrich, spacing=60000, sigma=150.0)
    out = np.empty((n, 4), dtype=np.int32)
    nt = max(1, min(8, os.cpu_count() or 1))
    step = (n + nt - 1) // nt
    ths = [threading.Thread(target=w.fragments, args=(i * step, min(n, (i + 1) * step) - i * step,
                                                      out[i * step:min(n, (i + 1) * step)]))
           for i in range(nt) if i * step < n]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


def test_hg38_full_size_properties(monkeypatch):
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 60 * (1 << 30):
        pytest.skip("needs ~60 GB of device memory")
    t = _fragments(N, 2001, 0.25)
    c = _fragments(N, 2002, 0.0)
    assert np.all(t[:, 1] >= 0) and np.all(c[:, 1] >= 0)
    api = capi.load_cuda()
    par = capi.make_params(p=0.01, keep_pileups=True)
    G = sum(HG38)

    def run(env, packed):
        for k in ("GR_FUSED", "GR_SB_MIN", "GR_FUSED_MIN"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = capi.Context(api, HG38, par)
        res = host.run_replicates(ctx, [(t, c)], packed=packed)
        return ctx, res

    ctx, res = run({}, False)
    st = res.sample_stats[0]
    tb = int(np.sum((np.minimum(t[:, 2], np.asarray(HG38)[t[:, 0]]) - t[:, 1]).astype(np.int64)))
    cb = int(np.sum((np.minimum(c[:, 2], np.asarray(HG38)[c[:, 0]]) - c[:, 1]).astype(np.int64)))
    assert st.frag_len == float(tb) and st.ctrl_frag == float(cb)
    assert np.float32(st.lambda_) == np.float32(tb / G)
    assert np.float32(st.factor) == np.float32(tb / cb)
    assert res.run_stats.genome_len == G
    total = 0
    n_iv = 0
    ends_by_chrom = []
    for ci, ln in enumerate(HG38):
        p = ctx.fetch(2, 0, ci)
        e64 = p.end.astype(np.int64)
        assert p.end[-1] == ln and np.all(np.diff(e64) > 0)
        assert np.all(p.val >= 0)
        if ci in (18, 20, 24):                         # the union check sorts: three chromosomes are enough
            ee, ce = ctx.fetch(0, 0, ci).end, ctx.fetch(1, 0, ci).end
            assert ee[-1] == ln and ce[-1] == ln
            assert np.array_equal(np.union1d(ee, ce), p.end)
        total += int(p.end[-1])
        n_iv += len(p.end)
        ends_by_chrom.append(p.end)
    assert total == G and n_iv == res.run_stats.n_intervals
    pk = res.peaks
    assert len(pk) > 10000
    assert np.all(pk["start"] >= 0) and np.all(pk["start"] < pk["end"])
    assert np.all(pk["end"] <= np.asarray(HG38)[pk["chrom"]])
    assert np.all(pk["summit"] < pk["end"] - pk["start"])
    key = pk["chrom"].astype(np.int64) * (1 << 32) + pk["start"]
    assert np.all(np.diff(key) > 0)
    assert np.all(pk["pval"] > par.min_pqval) and np.all(pk["auc"] >= par.min_auc)
    # peak boundaries are interval boundaries
    for ci in (0, 7, 24):
        sel = pk[pk["chrom"] == ci]
        assert np.all(np.isin(sel["end"].astype(np.uint32), ends_by_chrom[ci]))
    del ctx

    # dense formulation, second fused run with packed records: same bytes
    ctx_d, res_d = run({"GR_FUSED": "0"}, False)
    assert res_d.peaks.tobytes() == pk.tobytes()
    for ci in (0, 12, 24):
        assert np.array_equal(ctx_d.fetch(2, 0, ci).end, ends_by_chrom[ci])
    del ctx_d
    ctx_p, res_p = run({}, True)
    assert res_p.peaks.tobytes() == pk.tobytes()
    assert res_p.sample_stats[0].frag_len == st.frag_len
    print("hg38 full size: %d intervals, %d peaks" % (n_iv, len(pk)))
