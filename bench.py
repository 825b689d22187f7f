#!/usr/bin/env python
"""bench.py -- Gbp p-value-scanned per second of the pileup -> p -> q -> peak path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one whole pass of the hot path over one synthetic batch: zero the
dense delta arrays, scatter the treatment intervals, integrate, same for the
control, lambda / scale factor, control sweep, breakpoint union, -log10 p, (BH),
peak scan, peaks back on the host.  At N = 1 the workload is BASELINE.json
configs[1]: hg38-sized genome (25 chromosomes, 3.09 Gbp), 50 M treatment + 50 M
control paired-end fragments, default peak calling (-p 0.01).  For N > 1 the same
genome is sharded by chromosome over the ranks (strong scaling; torchrun, NCCL).

`value`  = genome bp / device time per step, interval records already in HBM.
`e2e`    = same through gr_push_packed() from PINNED HOST buffers (H2D copies
           inside the timed region) and peak records read back to the host.
`roofline` is for the dominant kernel, the per-base dense scan (k_scan_stream):
           4 B per delta cell per launch / mean launch time (CUDA events on the
           library's stream) against the measured HBM copy bandwidth.  Its
           companion `scan_place` (moves the breaks to their final rank, 16 B per
           interval) is timed as a stage of its own and quoted next to it.
`cpu_baseline` / --impl reference: the UNMODIFIED reference binary
           (oracle/_ref/Genrich, built by `make -C oracle ref` where the sources
           are) on the SAM view of a bounded sample of the same workload (8 x 50 Mbp,
           6.5 M + 6.5 M fragments: the workload's depth; ~19 s per run), 1 host
           core (the reference is single-threaded, README.md:535).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
        138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
        83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]

WORKLOADS = {
    # BASELINE.json configs[1]
    "hg38_chip_50M_50M": dict(chrom_len=HG38, nt=50_000_000, nc=50_000_000, q=None, p=0.01, atac=False,
                              spacing=60000, sigma=150.0, enrich=0.25),
    # configs[2]: ATAC mode, 100 M fragments, -q 0.05
    "hg38_atac_100M_q": dict(chrom_len=HG38, nt=100_000_000, nc=0, q=0.05, p=None, atac=True,
                             spacing=60000, sigma=60.0, enrich=0.3),
    # quick functional run
    "mini": dict(chrom_len=[60_000_000, 40_000_000, 20_000_000], nt=2_000_000, nc=2_000_000, q=None, p=0.01,
                 atac=False, spacing=40000, sigma=100.0, enrich=0.3),
}
# bounded CPU sample (same generator, same depth as the workload: 50 M fragments x 400 Mbp / 3.09 Gbp): ~19 s per run
SAMPLE = dict(chrom_len=[50_000_000] * 8, nt=6_500_000, nc=6_500_000)


# Kernel variants that exist behind environment knobs but are not the default because they have not
# been timed on a B200 yet (written where no GPU was at hand and checked on the CPU by tests/emu).
# The default run times each of them in a CHILD process (a fault or a hang there cannot take the
# headline with it), asserts that the peaks are the default path's byte for byte, and reports
# ms per step under "variants" -- information for the next round, never part of `value` / `e2e`.
VARIANTS = {           # simplest first: a faulting kernel poisons the child's context for everything after it
    "rm_per8": {"GR_RM_PER": "8"},
    "ur_groups2": {"GR_UR_GROUPS": "2"},
    "ur_groups4": {"GR_UR_GROUPS": "4"},
    "cl_tiles4": {"GR_CL_TILES": "4"},
    "ue_warp": {"GR_UE_WARP": "1"},
    "ue_pair": {"GR_UE_PAIR": "1"},
    "rank512": {"GR_FUSED_RANK": "1"},
    "rank1024": {"GR_FUSED_RANK": "1", "GR_FR_CAP": "1024"},
    "rank512_cps7": {"GR_FUSED_RANK": "1", "GR_FR_CPS": "7"},
    "rank512_pf4": {"GR_FUSED_RANK": "1", "GR_FR_PF": "4"},
    "p2": {"GR_FB_P2": "1"},
    "rank512_slots": {"GR_FUSED_RANK": "1", "GR_FB_SLOTS": "1"},
    "rank512_p2": {"GR_FUSED_RANK": "1", "GR_FB_P2": "1"},
    "all": {"GR_FUSED_RANK": "1", "GR_FB_SLOTS": "1", "GR_UE_WARP": "1", "GR_UR_GROUPS": "4", "GR_CL_TILES": "4"},
    "all_p2": {"GR_FUSED_RANK": "1", "GR_FB_P2": "1", "GR_UE_WARP": "1", "GR_UR_GROUPS": "4", "GR_CL_TILES": "4"},
    "all_p2_pair": {"GR_FUSED_RANK": "1", "GR_FB_P2": "1", "GR_UE_PAIR": "1", "GR_UR_GROUPS": "4", "GR_CL_TILES": "4"},
}


def run_variant_probe(a, timeout=300):
    cmd = [sys.executable, os.path.abspath(__file__), "--variant-probe", "--steps", "3", "--workload", a.workload]
    def last_json(text):
        for line in reversed((text or "").strip().split("\n")):
            if line.startswith("{"):
                try:
                    return json.loads(line)
                except ValueError:
                    continue
        return None
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired as e:
        # the child prints the results so far after every variant: a variant that hangs costs only itself
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else e.stdout
        got = last_json(out) or {}
        got["error"] = "variant probe timed out after %d s (results so far kept)" % timeout
        return got
    except Exception as e:
        return {"error": repr(e)[:300]}
    got = last_json(p.stdout)
    if got is not None:
        if p.returncode:
            got["error"] = "variant probe exit %d: %s" % (p.returncode, (p.stderr or "")[-300:])
        return got
    return {"error": "variant probe exit %d: %s" % (p.returncode, (p.stderr or p.stdout)[-300:])}


# dram__bytes_read.sum + dram__bytes_write.sum of one k_scan_stream launch (bytes), from the
# committed ncu capture of exactly this command; null for configurations that were not captured
NCU_TRAFFIC = {("hg38_chip_50M_50M", 1, "k_scan_stream"): 12.36e9 + 1.10e9,      # built array: nothing to clear behind
               ("hg38_chip_50M_50M", 1, "k_fb_scan"): 0.225e9 + 1.051e9}


def gen_fragments(chrom_len, n, seed, enrich, spacing, sigma, threads=8):
    from genrich_b200.synth import Workload
    w = Workload(chrom_len, n, seed, enrich=enrich, spacing=spacing, sigma=sigma)
    out = np.empty((n, 4), dtype=np.int32)
    step = (n + threads - 1) // threads
    ths = []
    for i in range(threads):
        a, b = i * step, min(n, (i + 1) * step)
        if a >= b:
            break
        t = threading.Thread(target=w.fragments, args=(a, b - a, out[a:b]))
        t.start()
        ths.append(t)
    for t in ths:
        t.join()
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).

    The poller is started well before the timed region (nvidia-smi needs ~1 s to come up) and
    every line carries a timestamp; stop(t0, t1) keeps the samples taken inside the timed window."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = float("nan")                    # unknown format: the sample still counts for the whole run
            try:
                rows.append((ts, float(c[2]), float(c[3]),
                             [n for n, v in zip(names, c[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if t0 is not None and t0 - 0.03 <= r[0] <= t1 + 0.03]
        window = "timed region"
        if len(inside) < 2:                          # clock skew / too short a run: everything since the start
            inside, window = rows, "whole run (fewer than 2 samples fell inside the timed region)"
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(r[2] for r in inside), "reasons": reasons,
                "samples": len(inside), "window": window}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------
def reference_sample_run(steps, warmup):
    """Time the unmodified reference on the SAM view of the bounded sample."""
    from genrich_b200.synth import Workload
    ref = os.path.join(ROOT, "oracle", "_ref", "Genrich")
    wl = WORKLOADS["hg38_chip_50M_50M"]
    L = SAMPLE["chrom_len"]
    td_ = tempfile.mkdtemp(prefix="grbench_")
    tp, cp, op = (os.path.join(td_, x) for x in ("t.sam", "c.sam", "o.np"))
    Workload(L, SAMPLE["nt"], 3001, enrich=wl["enrich"], spacing=wl["spacing"], sigma=wl["sigma"]).write_sam(tp)
    Workload(L, SAMPLE["nc"], 3002, enrich=0.0).write_sam(cp)
    G = sum(L)
    kind = "reference"
    cmd = [ref, "-t", tp, "-c", cp, "-o", op, "-p", "0.01"]
    if not os.path.exists(ref):
        raise SystemExit("oracle/_ref/Genrich missing: run `make -C oracle ref` where /root/reference exists")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    npk = sum(1 for _ in open(op))
    for f in (tp, cp, op):
        os.unlink(f)
    os.rmdir(td_)
    sec = sum(times) / len(times)
    return {"value": G / 1e9 / sec, "unit": "Gbp/s", "cores": 1, "kind": kind,
            "sample": "%d chrom x %d bp, %d + %d fragments (same generator), whole program incl. SAM parse, %.2f s/run, %d peaks"
                      % (len(L), L[0], SAMPLE["nt"], SAMPLE["nc"], sec, npk),
            "host_cores_available": os.cpu_count()}, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="hg38_chip_50M_50M")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-formulation (GR_FUSED=0) comparison pass")
    ap.add_argument("--no-variants", action="store_true", help="skip the child process that times the non-default kernel variants")
    ap.add_argument("--variant-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-pack6", action="store_true", help="e2e arm with 8-byte records even where 6-byte records fit")
    ap.add_argument("--prefetch-depth", type=int, default=2,
                    help="e2e arm: samples sent ahead of their push (2: both samples of the next step, 1: the next sample)")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[a.workload]
    L = wl["chrom_len"]
    G = sum(L)

    if a.impl == "reference":
        if rank != 0:
            return
        cb, sec = reference_sample_run(a.steps, 1)
        line = {"impl": "reference", "metric": "Gbp p-value-scanned/sec", "value": cb["value"], "unit": "Gbp/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": 1, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f32/f64",
                "data": "synthetic", "config": {"workload": a.workload, "sample": cb["sample"]},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as td
    from genrich_b200 import capi, host
    from genrich_b200.dist import ShardedEngine

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    if rank == 0 and not a.variant_probe:
        sampler.start()                                # comes up while the workload is generated
    host_group = None
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
        host_group = td.new_group(backend="gloo")      # host-resident scalars and peak records
    api = capi.load_cuda()
    par = capi.make_params(p=wl["p"], q=wl["q"])
    eng = ShardedEngine(api, L, par, dev, host_group=host_group)
    ctx = eng.ctx

    # synthetic interval records of this rank's chromosomes, on the host (pinned) and in HBM
    def make(n, seed, enrich):
        if n == 0:
            return None, None, None
        fr = gen_fragments(L, n, seed, enrich, wl["spacing"], wl["sigma"])
        iv = eng.route(host.fragments_to_intervals(fr, atac=wl["atac"]))
        # the producer side of the C-ABI hands over 8-byte GR_PACK records (gr_push_packed) ...
        packed, rest = host.pack_records(iv)
        assert rest.shape[0] == 0, "synthetic workload has records that do not pack"
        pinned = torch.from_numpy(packed.view(np.int64)).pin_memory()
        # ... or, where the context's layout fits 32 bits, 6-byte GR_PACK6 records (gr_push_packed6):
        # the e2e arm is bound by the host -> device copy, 25 % fewer bytes
        p6 = None
        if layout6 is not None and not a.no_pack6:
            r6, rest6 = host.pack6_records(iv, layout6, L)
            if rest6.shape[0] == 0:
                p6 = torch.from_numpy(r6.view(np.int16).reshape(-1)).pin_memory()
        return pinned, pinned.to(dev), p6
    # N > 1 keeps the record format the multi-GPU runs of this round were validated with (8-byte words,
    # one sample ahead) unless asked otherwise; the 6-byte path itself is rank-agnostic
    layout6 = ctx.pack6_layout() if (world == 1 or os.environ.get("GR_BENCH_PACK6_MULTI")) else None
    t_host, t_dev, t_h6 = make(wl["nt"], 2001, wl["enrich"])
    c_host, c_dev, c_h6 = make(wl["nc"], 2002, 0.0)
    use6 = t_h6 is not None and (c_host is None or c_h6 is not None)
    n_t = t_host.shape[0]
    n_c = c_host.shape[0] if c_host is not None else 0
    torch.cuda.synchronize()

    def step(from_host, eng=eng):
        ctx = eng.ctx
        ctx.reset()
        eng.saved_any[:] = False
        eng.sample_stats.clear()
        ahead = None
        if from_host and use6:
            # Both samples of the NEXT step are sent while this step computes (the library's two
            # prefetch slots; a slot is refilled as soon as the pileup that read it is done), so
            # in steady state every step's 0.6 GB of input travels under the previous step's kernels.
            def pe(c):
                c.push_packed6_ptr(t_h6.data_ptr(), n_t)
                if n_c and a.prefetch_depth < 2:
                    c.prefetch_packed6_ptr(c_h6.data_ptr(), n_c)

            def pc_(c):
                c.push_packed6_ptr(c_h6.data_ptr(), n_c)
                if a.prefetch_depth < 2:
                    c.prefetch_packed6_ptr(t_h6.data_ptr(), n_t)
            pc = pc_ if n_c else None
            if a.prefetch_depth >= 2 or not n_c:
                def ahead(c):
                    c.prefetch_packed6_ptr(t_h6.data_ptr(), n_t)
                    if n_c:
                        c.prefetch_packed6_ptr(c_h6.data_ptr(), n_c)
        elif from_host:
            # host (pinned) -> device copies are inside the timed region; the control sample is
            # sent while the treatment sample is being integrated, and the next step's treatment
            # sample while this step's peaks are called (gr_prefetch_intervals)
            def pe(c):
                c.push_packed_ptr(t_host.data_ptr(), n_t)
                if n_c:
                    c.prefetch_packed_ptr(c_host.data_ptr(), n_c)

            def pc_(c):
                c.push_packed_ptr(c_host.data_ptr(), n_c)
                c.prefetch_packed_ptr(t_host.data_ptr(), n_t)
            pc = pc_ if n_c else None
        else:
            pe = lambda c: c.push_packed_ptr(t_dev.data_ptr(), n_t)
            pc = (lambda c: c.push_packed_ptr(c_dev.data_ptr(), n_c)) if n_c else None
        eng.replicate(pe, pc, want_stats=False)
        if ahead is not None:
            ahead(eng.ctx)
        return eng.call_peaks()

    trace = {}
    if os.environ.get("GR_BENCH_TRACE"):
        # host time of every library call of a step (debugging aid: where the host keeps the GPU waiting)
        def wrap(obj, name, key=None):
            f = getattr(obj, name)
            name = key or name

            def g(*a_, **k_):
                t0 = time.perf_counter()
                r = f(*a_, **k_)
                trace[name] = trace.get(name, 0.0) + time.perf_counter() - t0
                return r
            setattr(obj, f.__name__, g)
        for nm in ("reset", "sample_begin", "push_packed_ptr", "prefetch_packed_ptr", "sample_pileup_async",
                   "replicate_finish_device", "pvalues_finalize", "call_peaks", "timer_start", "timer_stop"):
            wrap(eng.ctx, nm)
        wrap(eng, "replicate", "eng.replicate (incl. the calls above)")
        wrap(eng, "call_peaks", "eng.call_peaks (incl. call_peaks)")

    def timed(from_host, steps, warmup, with_stages=False, eng=eng):
        ctx = eng.ctx
        for _ in range(warmup):
            step(from_host, eng)
        eng.t_acc.clear()
        if with_stages:
            ctx.timing(True)
            ctx.timing_reset()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        l0 = ctx.kernel_launches()
        dev_ms = []
        w0 = time.perf_counter()
        for _ in range(steps):
            ctx.timer_start()
            peaks, rs = step(from_host, eng)
            dev_ms.append(ctx.timer_stop())
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        wall = (time.perf_counter() - w0) * 1e3 / steps
        ms = sum(dev_ms) / len(dev_ms)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        stages = ctx.timing_get() if with_stages else None
        if with_stages:
            ctx.timing(False)
        return float(t[0]), float(t[1]), ctx.kernel_launches() - l0, peaks, rs, stages

    if a.variant_probe:
        # child of the default run: every variant against the default path, same records, fresh context each
        _, _, _, base_peaks, _, _ = timed(False, 2, 3)
        out = {}
        for name, env in VARIANTS.items():
            eng_v = None
            try:
                os.environ.update(env)
                eng_v = ShardedEngine(api, L, par, dev, host_group=None)
                ms_v, _, launches_v, peaks_v, _, st_v = timed(False, a.steps, 3, with_stages=True, eng=eng_v)
                out[name] = {"env": env, "ms_per_step_with_stage_events": round(ms_v, 4), "launches": int(launches_v),
                             "peaks_identical": bool(peaks_v.tobytes() == base_peaks.tobytes()), "peaks": int(len(peaks_v)),
                             "stage_ms_per_step": {k: round(v[0] / a.steps, 4) for k, v in
                                                   sorted(st_v.items(), key=lambda kv: -kv[1][0])[:6]}}
            except Exception as e:                     # a variant that fails says so; the others still run
                out[name] = {"env": env, "error": repr(e)[:300]}
            finally:
                for k in env:
                    os.environ.pop(k, None)
                if eng_v is not None:
                    try:
                        eng_v.ctx.close()              # its device buffers go back before the next variant allocates
                    except Exception:
                        pass
            print(json.dumps(out), flush=True)         # cumulative: the parent keeps the last complete line
        return

    t_w0 = time.time()
    # headline: K steps, nothing but the step itself in the stream; the per-stage CUDA events (two
    # per stage, ~20 stages per step) are recorded in a separate short pass of the same step
    ms_dev, wall_dev, launches, peaks, rs, _ = timed(False, a.steps, a.warmup)
    if trace:
        print("host ms per step by call (device arm, %d steps incl. warm-up): %s" % (
            a.steps + a.warmup, {k: round(v * 1e3 / (a.steps + a.warmup), 3) for k, v in trace.items()}),
            file=sys.stderr, flush=True)
        trace.clear()
    ms_e2e, wall_e2e, _, peaks2, _, _ = timed(True, a.steps, 1)
    # rank 0's host time by phase of the e2e arm (where the host waits for the device it is device time too)
    host_phase = {k: round(v * 1e3 / a.steps, 4) for k, v in eng.t_acc.items()}
    st_steps = max(2, a.steps // 2)
    ms_staged, _, _, peaks3, _, stages = timed(False, st_steps, 1, with_stages=True)
    assert peaks3.tobytes() == peaks.tobytes()
    clocks = sampler.stop(t_w0, time.time()) if rank == 0 else None
    # the same per-base pass in its dense formulation (delta array in HBM: bucketed build, then
    # k_scan_stream reads 4 B per cell): the kernel the HBM-read roofline is literally about
    dense = None
    if world == 1 and not a.no_dense:
        os.environ["GR_FUSED"] = "0"
        eng_d = ShardedEngine(api, L, par, dev, host_group=None)
        del os.environ["GR_FUSED"]
        ms_d, _, _, peaks_d, _, st_d = timed(False, st_steps, 3, with_stages=True, eng=eng_d)
        assert peaks_d.tobytes() == peaks.tobytes(), "dense and fused formulations disagree"
        dense = (ms_d, st_d)
        del eng_d

    variants = None
    if world == 1 and not a.no_variants and a.workload != "mini":
        variants = run_variant_probe(a)                # a child process; this one is idle meanwhile

    if eng.debug:
        print("rank %d host-side ms per step (e2e arm): %s" % (rank, {k: round(v * 1e3 / a.steps, 3) for k, v in eng.t_acc.items()}),
              file=sys.stderr, flush=True)
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    n_samples = 2 if n_c else 1
    peak_gbs, peak_src = measured_peak_gbs()
    fused = "fused_scan" in stages
    scan_kernel = "k_fb_scan" if fused else "k_scan_stream"
    scan_ms, scan_launches, _ = stages.get("fused_scan" if fused else "dense_scan", (0.0, 0, 0))
    per_launch_ms = scan_ms / max(scan_launches, 1)          # mean over every launch of the staged pass
    place_ms, place_launches, _ = stages.get("scan_place", (0.0, 0, 0))
    cells = ctx_cells = sum((int(l) + 1 + 8191) // 8192 * 8192 for l, o in zip(L, eng.owned) if o)
    achieved = 4.0 * cells / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms else 0.0
    dense_obj = None
    if dense is not None:
        d_ms, d_n, _ = dense[1].get("dense_scan", (0.0, 0, 0))
        b_ms, b_n, _ = dense[1].get("build", (0.0, 0, 0))
        d_per = d_ms / max(d_n, 1)
        d_ach = 4.0 * cells / (d_per * 1e-3) / 1e9 if d_per else 0.0
        dense_obj = {"kernel": "k_scan_stream", "achieved": d_ach, "peak": peak_gbs, "unit": "GB/s",
                     "frac": d_ach / peak_gbs if peak_gbs else None, "ms_per_launch": d_per,
                     "build_ms_per_launch": b_ms / max(b_n, 1), "ms_per_step": dense[0],
                     "traffic": NCU_TRAFFIC.get((a.workload, world, "k_scan_stream")),
                     "note": "GR_FUSED=0: the delta array is written to HBM by k_sb_build and read back by "
                             "k_scan_stream (4 B per cell each way); same peaks, bit for bit"}
    stage_ms = {k: round(v[0] / st_steps, 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])}
    stage_ms["_step_with_stage_events"] = round(ms_staged, 4)
    line = {
        "metric": "Gbp p-value-scanned/sec", "value": G / 1e9 / (ms_dev * 1e-3), "unit": "Gbp/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 deltas, f32 pileups, f64 -> f32 -log10 p", "data": "synthetic",
        "config": {"workload": a.workload, "genome_bp": G, "chromosomes": len(L), "treatment_fragments": wl["nt"],
                   "control_fragments": wl["nc"], "threshold": "-q %g" % wl["q"] if wl["q"] else "-p %g" % wl["p"],
                   "atac": wl["atac"], "sharding": "chromosomes over %d rank(s), LPT" % world,
                   "l2": "inputs (%.1f GB dense delta array per sample) far exceed the 126 MB L2" % (4e-9 * cells),
                   "peaks": int(len(peaks)), "intervals_rank0": int(rs.n_intervals)},
        "e2e": {"value": G / 1e9 / (ms_e2e * 1e-3), "unit": "Gbp/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int((6 if use6 else 8) * (n_t + n_c)), "d2h_bytes_per_step": int(peaks2.nbytes + 512),
                "record_format": "GR_PACK6 (6 B per record, expanded on the device)" if use6 else "GR_PACK (8 B per record)",
                "wall_ms_per_step": wall_e2e,
                "peaks_identical_to_device_arm": bool(peaks2.tobytes() == peaks.tobytes())},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall_dev,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs if peak_gbs else None,
                     # ncu --set full, hg38 workload, 1 GPU (profiles/r01_scan_stream_ncu.txt): dram read + write per launch
                     "traffic": NCU_TRAFFIC.get((a.workload, world, scan_kernel)), "peak_source": peak_src,
                     "bytes_per_launch": 4 * cells, "ms_per_launch": per_launch_ms,
                     "companion_scan_place_ms_per_launch": place_ms / max(place_launches, 1),
                     "launches_per_step": scan_launches / st_steps, "samples_scanned_per_step": n_samples,
                     "note": ("the delta cells of k_fb_scan live in shared memory only: `achieved` divides the ALGORITHMIC "
                              "bytes of the per-base pass (SURVEY 8d: 4 B per base per sample array) by the launch time, "
                              "`traffic` is what the kernel really moves through DRAM (bucket entries in, breaks and "
                              "bitmap out); frac > 1 = faster than any kernel that reads the array from HBM could be"
                              if fused else "4 B per delta cell read from HBM"),
                     "dense_formulation": dense_obj},
        "stage_ms_per_step": stage_ms,
        "variants": variants,
        "host_phase_ms_per_step_e2e": host_phase,
    }
    if world > 1:
        td.destroy_process_group()                 # nothing below involves the other ranks
    if world == 1 and not a.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Genrich")):
        cb, _ = reference_sample_run(1, 0)         # the contract: on rank 0 at N = 1 only
        line["cpu_baseline"] = cb
    else:
        line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": 1, "kind": "reference",
                                "sample": "skipped (N > 1, --no-cpu-baseline or oracle/_ref missing)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
