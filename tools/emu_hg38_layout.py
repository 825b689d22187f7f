#!/usr/bin/env python
"""Development aid (CPU, ~6 minutes, 25 GB of RAM): the knob-gated kernel variants against the default path on
the hg38-SIZED layout (25 chromosomes, 3.09 G cells: block numbers, bitmap words and cell offsets beyond
32 bits of bytes) with a sparse sample (400 k + 400 k fragments, 6-byte records), through the CPU-emulated
library (make -C tests/emu _build/libgenrich_emu.so first).  Run once after touching index arithmetic."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genrich_b200 import capi, host
from genrich_b200.synth import Workload
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
        138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
        83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
api = capi.Api(os.path.join(ROOT, "tests", "emu", "_build", "libgenrich_emu.so"), "gr_")
n = 400000
t = Workload(HG38, n, 11, enrich=0.6, spacing=3000000, sigma=150.0).fragments()
c = Workload(HG38, n, 12, enrich=0.0).fragments()
par = capi.make_params(p=0.01, min_auc=5.0)
modes = {"default": {}, "all": {"GR_FUSED_RANK": "1", "GR_FB_SLOTS": "1", "GR_UE_WARP": "1", "GR_UR_GROUPS": "4", "GR_CL_TILES": "4"},
         "all_p2": {"GR_FUSED_RANK": "1", "GR_FR_CAP": "1024", "GR_FB_P2": "1", "GR_UE_WARP": "1", "GR_UR_GROUPS": "2", "GR_CL_TILES": "4"}}
res = {}
for name, env in modes.items():
    os.environ.update({"GR_FUSED": "1", "GR_FUSED_MIN": "1"}); os.environ.update(env)
    t0 = time.time()
    ctx = capi.Context(api, HG38, par)
    r = host.run_replicates(ctx, [(t, c)], packed=6)
    st = r.sample_stats[0]
    res[name] = (r.peaks.tobytes(), st.frag_len, st.ctrl_frag, st.n_expt, st.n_ctrl, st.n_pval, st.lambda_, st.factor)
    print(name, "%.1fs" % (time.time() - t0), len(r.peaks), "peaks", st.n_expt, st.n_ctrl, st.n_pval, st.frag_len, flush=True)
    ctx.close()
    for k in env: os.environ.pop(k)
assert res["all"] == res["default"] and res["all_p2"] == res["default"]
assert res["default"][1] == float(np.sum((t[:, 2] - t[:, 1]).astype(np.int64)))
print("hg38-sized layout: variants identical to the default path; exact fragment sum")
