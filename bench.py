#!/usr/bin/env python
"""bench.py -- Gbp p-value-scanned per second of the pileup -> p -> q -> peak path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one whole pass of the hot path over one synthetic batch: for every replicate the
treatment records are bucketed and integrated, same for its control, lambda / scale factor,
control sweep, breakpoint union, -log10 p; then (Fisher combine), (Benjamini-Hochberg incl. the
histogram all-gather at N > 1), the peak scan, and the peaks back on the host.

Workloads (BASELINE.json `configs`; --workload):
  hg38_chip_50M_50M   configs[1]  hg38-sized (25 chromosomes, 3.09 Gbp), 50 M + 50 M fragments, -p 0.01
                                  -- the configuration the metric is quoted on; the default at every N
  hg38_atac_100M_q    configs[2]  ATAC (-j -d 100), 100 M fragments, -q 0.05
  hg38_fisher3        configs[3]  3 replicates x 50 M + 1 control, Fisher's method, -p 0.01
  g10_multimap_1B_q   configs[4]  10 Gbp (40 x 250 Mbp), 1 B fragments, 30 % multimapped (-s 20), -q 0.05
                                  (needs >= 8 GPUs' worth of memory: run it with --gpus 8)
  g10_shard_125M_q                one rank's share of configs[4] (5 x 250 Mbp, 125 M fragments): its 1-GPU twin
  mini                            quick functional run
For N > 1 the same genome is sharded by chromosome over the ranks (strong scaling; torchrun, NCCL).

`value`  = genome bp / device time per step, interval records already in HBM.
`e2e`    = same through gr_push_packed6 / gr_push_packed from PINNED HOST buffers (H2D copies inside
           the timed region) with the peak records read back to the host.
`roofline` is for the per-base pass (k_fr_scan or k_fb_scan, whichever the sample chose on the device --
           `roofline.scan_form`): ALGORITHMIC bytes (4 B per delta cell per sample
           array) / mean launch time, next to what the kernel and the whole step really
           move through DRAM (`traffic`, `frac_dram`, `step`), from the committed ncu pass
           profiles/r02_dram_by_stage.json, and to the dense formulation (k_scan_stream).
`parity` = in-run gate: bounded samples pushed through the GPU and compared with the narrowPeak
           files the UNMODIFIED reference wrote for the same SAM view in this very run.
`cpu_baseline` / --impl reference: oracle/_ref/Genrich (built by `make -C oracle ref` where the
           sources are) on the SAM view of a bounded, depth-matched sample of the workload, 1 host core
           (the reference is single-threaded, README.md:535); `hot_path` = the share of that time spent
           in the functions of the path (gprof, -pg build), i.e. without SAM parsing.
"""
import argparse
import hashlib
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
G10 = [250_000_000] * 40

# reps: per replicate (treatment fragments, control fragments or 0).  sample: the bounded CPU sample of
# the same shape and depth (same generator) that the reference is timed and the parity gate is run on.
WORKLOADS = {
    "hg38_chip_50M_50M": dict(chrom_len=HG38, reps=[(50_000_000, 50_000_000)], q=None, p=0.01, atac=False,
                              spacing=60000, sigma=150.0, enrich=0.25, multimap=0.0,
                              sample=dict(chrom_len=[50_000_000] * 8, reps=[(6_500_000, 6_500_000)])),
    # ATAC without a control: lambda ~ 6.5, and the log-normal tail at mu < 7 is heavy (a 50-fold pileup is only
    # -log10 p = 5.9 < log10 G: "All q-values are 1") -- few, tall, narrow sites so that the reference calls peaks under -q
    "hg38_atac_100M_q": dict(chrom_len=HG38, reps=[(100_000_000, 0)], q=0.05, p=None, atac=True,
                             spacing=1_000_000, sigma=40.0, enrich=0.4, multimap=0.0,
                             sample=dict(chrom_len=[50_000_000] * 4, reps=[(6_500_000, 0)])),
    "hg38_fisher3": dict(chrom_len=HG38, reps=[(50_000_000, 50_000_000), (50_000_000, 0), (50_000_000, 0)],
                         q=None, p=0.01, atac=False, spacing=60000, sigma=150.0, enrich=0.25, multimap=0.0,
                         sample=dict(chrom_len=[50_000_000] * 4, reps=[(3_250_000, 3_250_000), (3_250_000, 0), (3_250_000, 0)])),
    "g10_multimap_1B_q": dict(chrom_len=G10, reps=[(1_000_000_000, 0)], q=0.05, p=None, atac=False,
                              spacing=60000, sigma=150.0, enrich=0.25, multimap=0.3,
                              sample=dict(chrom_len=[50_000_000], reps=[(5_000_000, 0)])),
    "g10_shard_125M_q": dict(chrom_len=G10[:5], reps=[(125_000_000, 0)], q=0.05, p=None, atac=False,
                             spacing=60000, sigma=150.0, enrich=0.25, multimap=0.3,
                             sample=dict(chrom_len=[50_000_000], reps=[(5_000_000, 0)])),
    "mini": dict(chrom_len=[60_000_000, 40_000_000, 20_000_000], reps=[(2_000_000, 2_000_000)], q=None, p=0.01,
                 atac=False, spacing=40000, sigma=100.0, enrich=0.3, multimap=0.0,
                 sample=dict(chrom_len=[20_000_000] * 2, reps=[(700_000, 700_000)])),
}
# what the default run pushes through BOTH the reference and the GPU (the in-run parity gate):
# the workload's own sample plus one -q and one Fisher + -q sample, so that BH and the combine are reference-checked too
GATE_EXTRA = {
    "atac_q": dict(chrom_len=[30_000_000] * 2, reps=[(2_000_000, 0)], q=0.05, p=None, atac=True,
                   spacing=500_000, sigma=40.0, enrich=0.4, multimap=0.0),
    "fisher_multimap_q": dict(chrom_len=[12_000_000], reps=[(350_000, 350_000), (350_000, 0), (350_000, 0)],
                              q=0.05, p=None, atac=False, spacing=60000, sigma=150.0, enrich=0.3, multimap=0.3),
}
HOT_FUNCS = ("saveInterval", "addFrac", "subFrac", "savePileupExpt", "savePileupCtrl", "savePileupNoCtrl", "saveLambda",
             "saveConst", "calcLambda", "calcFactor", "updateVal", "getVal", "savePval", "countIntervals", "calcPval",
             "plnorm", "pnorm", "do_del", "combinePval", "countIntervals2", "multPval", "pchisq", "pgamma",
             "pgamma_smallx", "pd_upper_series", "pd_lower_series", "dpois", "stirlerr", "bd0", "computeQval",
             "hashPval", "recordPval", "jenkins_one_at_a_time_hash", "collectPval", "saveQval", "quickSort",
             "partition", "lookup", "findPeaks", "callPeaks", "updatePeak", "checkPeak", "resetVars")


# --------------------------------------------------------------------------------------
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


def bind_to_gpu_node(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned staging buffers are
    allocated (first touch then puts them on that node): eight ranks pulling records from host memory through the
    wrong socket is what held the round-1 e2e arm at 21 GB/s per GPU.  Best effort: any failure leaves things as they are."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                # sysfs uses a 4-digit domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"node": node, "cpus": len(allowed)}
    except (OSError, ValueError, subprocess.SubprocessError):
        pass
    return None


def peaks_sha256(peaks):
    """Hash of what the parity bar makes bit-exact: chromosome, start, end, summit of every peak, in order."""
    h = hashlib.sha256()
    for f in ("chrom", "start", "end", "summit"):
        h.update(np.ascontiguousarray(peaks[f]).astype(np.int64).tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------------------
# workload generation: per chromosome (own seed), so that a rank generates only what it owns and
# every N sees the same genome-wide data; in chunks, so that the host footprint stays small at 1 B fragments
CHUNK = 4_000_000


def chrom_share(chrom_len, n):
    """fragments per chromosome, proportional to length (largest remainder: sums to n exactly)"""
    L = np.asarray(chrom_len, dtype=np.float64)
    raw = L / L.sum() * n
    base = np.floor(raw).astype(np.int64)
    rest = int(n - base.sum())
    if rest:
        base[np.argsort(-(raw - base), kind="stable")[:rest]] += 1
    return base


def gen_tasks(wl, n, seed, enrich, owned):
    share = chrom_share(wl["chrom_len"], n)
    tasks = []
    for c, nc in enumerate(share):
        if not owned[c]:
            continue
        chunk = CHUNK // 4 if wl["multimap"] else CHUNK       # the generator reserves 10 records per multimapped fragment
        for first in range(0, int(nc), chunk):
            tasks.append((c, first, min(chunk, int(nc) - first), seed * 1000 + c, enrich))
    return tasks


def run_tasks(wl, tasks, consume, threads=8):
    """consume(idx, records int32 (m, 4)) is called from worker threads, once per task"""
    from genrich_b200.synth import Workload
    lock = threading.Lock()
    it = iter(enumerate(tasks))

    def work():
        while True:
            with lock:
                nx = next(it, None)
            if nx is None:
                return
            i, (c, first, m, seed, enrich) = nx
            w = Workload([wl["chrom_len"][c]], first + m, seed, enrich=enrich, spacing=wl["spacing"], sigma=wl["sigma"],
                         multimap=wl["multimap"])
            fr = w.fragments(first, m)               # ctypes releases the GIL: the workers run in parallel
            fr[:, 0] = c
            consume(i, fr)
    ths = [threading.Thread(target=work) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()


class Sample:
    """One input file's worth of interval records: 8-byte words in HBM (device arm), 6- or 8-byte records in
    pinned host memory (e2e arm)."""

    def __init__(self, torch, host_mod, eng, wl, n, seed, enrich, dev, use6, threads, tag):
        layout6 = eng.ctx.pack6_layout() if use6 else None
        L = wl["chrom_len"]
        # GR_BENCH_CACHE=<dir>: packed records are kept there between runs of one session (the generator is
        # deterministic; a profiling session runs bench.py many times on the same workload)
        cache = os.environ.get("GR_BENCH_CACHE")
        f8 = os.path.join(cache, tag + ".p8.npy") if cache else None
        f6 = os.path.join(cache, tag + ".p6.npy") if cache else None
        if cache and os.path.exists(f8) and (layout6 is None or os.path.exists(f6)):
            p8 = [np.load(f8)]
            p6 = [np.load(f6)] if layout6 is not None else []
        else:
            tasks = gen_tasks(wl, n, seed, enrich, eng.owned)
            p8 = [None] * len(tasks)
            p6 = [None] * len(tasks)

            def consume(i, fr):
                iv = host_mod.fragments_to_intervals(fr, atac=wl["atac"])
                if wl["atac"]:
                    # cut-site intervals reach past the chromosome ends: clamped here as saveInterval does
                    # (2522-2544) -- the packed record forms hold no negative start
                    np.clip(iv[:, 1], 0, None, out=iv[:, 1])
                    np.minimum(iv[:, 2], np.asarray(L, dtype=np.int32)[iv[:, 0]], out=iv[:, 2])
                a, rest = host_mod.pack_records(iv)
                assert rest.shape[0] == 0, "synthetic workload has records that do not pack"
                p8[i] = a
                if layout6 is not None:
                    b, rest6 = host_mod.pack6_records(iv, layout6, L)
                    assert rest6.shape[0] == 0, "synthetic workload has records that do not fit 6 bytes"
                    p6[i] = b.reshape(-1)
            run_tasks(wl, tasks, consume, threads)
            if cache:
                os.makedirs(cache, exist_ok=True)
                np.save(f8, np.concatenate(p8) if p8 else np.empty(0, np.uint64))
                if layout6 is not None:
                    np.save(f6, np.concatenate(p6) if p6 else np.empty(0, np.uint16))
        self.n = int(sum(len(a) for a in p8))
        self.fmt = 6 if layout6 is not None else 8
        self.dev = torch.empty(max(self.n, 1), dtype=torch.int64, device=dev)
        at = 0
        for a in p8:                                 # chunk by chunk: no second full-size host copy
            self.dev[at:at + len(a)].copy_(torch.from_numpy(a.view(np.int64)))
            at += len(a)
        if self.fmt == 6:
            self.host = torch.empty(max(3 * self.n, 1), dtype=torch.int16, pin_memory=True)
            at = 0
            for b in p6:
                self.host[at:at + len(b)].copy_(torch.from_numpy(b.view(np.int16)))
                at += len(b)
        else:
            self.host = torch.empty(max(self.n, 1), dtype=torch.int64, pin_memory=True)
            at = 0
            for a in p8:
                self.host[at:at + len(a)].copy_(torch.from_numpy(a.view(np.int64)))
                at += len(a)
        self.host_bytes = self.fmt * self.n

    def push_dev(self, c):
        c.push_packed_ptr(self.dev.data_ptr(), self.n)

    def push_host(self, c):
        (c.push_packed6_ptr if self.fmt == 6 else c.push_packed_ptr)(self.host.data_ptr(), self.n)

    def prefetch(self, c):
        (c.prefetch_packed6_ptr if self.fmt == 6 else c.prefetch_packed_ptr)(self.host.data_ptr(), self.n)


class Feeder:
    """e2e arm: the samples of a step in push order, cyclic.  The library has two prefetch slots; a slot
    is free again once the pileup that read it has been enqueued.  top_up() is called right after pileups
    were enqueued (no slot in use), so len(ahead) == slots whose copy is under way or done."""

    def __init__(self, order, depth):
        self.order, self.depth = order, depth
        self.seq = 0                                 # sequence number of the next push
        self.ahead = []                              # sequence numbers prefetched and not pushed yet

    def top_up(self, c):
        while len(self.ahead) < self.depth:
            k = (self.ahead[-1] if self.ahead else self.seq - 1) + 1
            s = self.order[k % len(self.order)]
            if s.n:
                s.prefetch(c)
            self.ahead.append(k)

    def push(self, c):
        self.top_up(c)                               # the previous sample's pileup is enqueued by now
        if self.ahead and self.ahead[0] == self.seq:
            self.ahead.pop(0)
        self.order[self.seq % len(self.order)].push_host(c)
        self.seq += 1


# --------------------------------------------------------------------------------------
# the unmodified reference on the SAM view of a bounded sample
def sample_workloads(spec, wl):
    """one classic multi-chromosome Workload per input file of the sample (SAM view and records agree)"""
    from genrich_b200.synth import Workload
    L = spec["chrom_len"]
    out = []
    for r, (nt, nc) in enumerate(spec["reps"]):
        t = Workload(L, nt, 3001 + 10 * r, enrich=wl["enrich"], spacing=wl["spacing"], sigma=wl["sigma"], multimap=wl["multimap"])
        c = Workload(L, nc, 3002 + 10 * r, enrich=0.0, multimap=wl["multimap"]) if nc else None
        out.append((t, c))
    return out


def ref_args(wl):
    a = ["-q", "%g" % wl["q"]] if wl["q"] else ["-p", "%g" % wl["p"]]
    if wl["atac"]:
        a += ["-j", "-d", "100"]
    if wl["multimap"]:
        a += ["-s", "20"]
    return a


def parse_narrowpeak(path):
    rows = []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        rows.append((int(f[0][3:]) - 1, int(f[1]), int(f[2]), int(f[9]), float(f[7]), float(f[8])))
    return rows


class RefSample:
    """SAM files of a bounded sample on disk + the reference's own runs on them."""

    def __init__(self, spec, wl, name):
        self.spec, self.wl, self.name = spec, wl, name
        self.dir = tempfile.mkdtemp(prefix="grbench_")
        self.wls = sample_workloads(spec, wl)
        self.tfiles, self.cfiles = [], []
        for r, (t, c) in enumerate(self.wls):
            tp = os.path.join(self.dir, "t%d.sam" % r)
            t.write_sam(tp)
            self.tfiles.append(tp)
            if c is not None:
                cp = os.path.join(self.dir, "c%d.sam" % r)
                c.write_sam(cp)
                self.cfiles.append(cp)
            else:
                self.cfiles.append("null")
        self.G = sum(spec["chrom_len"])
        self.out = os.path.join(self.dir, "ref.narrowPeak")

    def cmd(self, binary, out):
        c = [binary, "-t", ",".join(self.tfiles), "-o", out] + ref_args(self.wl)
        if any(x != "null" for x in self.cfiles):
            c += ["-c", ",".join(self.cfiles)]
        return c

    def run_reference(self, runs=1):
        ref = os.path.join(ROOT, "oracle", "_ref", "Genrich")
        if not os.path.exists(ref):
            raise SystemExit("oracle/_ref/Genrich missing: run `make -C oracle ref` where /root/reference exists")
        times = []
        for _ in range(runs):
            t0 = time.perf_counter()
            subprocess.check_call(self.cmd(ref, self.out), stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t0)
        return times

    def hot_path(self, wall):
        """share of the reference's time spent in the functions of the path: -pg build, gprof self time of the
        path's functions, scaled by wall / wall_pg (the instrumented run is slower)"""
        pg = os.path.join(ROOT, "oracle", "_ref", "Genrich_pg")
        if not os.path.exists(pg):
            return None
        try:
            t0 = time.perf_counter()
            subprocess.check_call(self.cmd(pg, os.path.join(self.dir, "pg.narrowPeak")), stderr=subprocess.DEVNULL, cwd=self.dir)
            wall_pg = time.perf_counter() - t0
            txt = subprocess.run(["gprof", "-b", "-p", pg, os.path.join(self.dir, "gmon.out")], capture_output=True,
                                 text=True, timeout=120).stdout
        except (OSError, subprocess.SubprocessError):
            return None
        hot = tot = 0.0
        for line in txt.split("\n"):
            f = line.split()
            if len(f) < 4:
                continue
            try:
                self_s = float(f[2])
                float(f[0])
            except ValueError:
                continue
            tot += self_s
            if f[-1] in HOT_FUNCS:
                hot += self_s
        if tot <= 0:
            return None
        t_hot = hot * wall / wall_pg
        return {"seconds": round(t_hot, 3), "share_of_wall": round(t_hot / wall, 4), "value": self.G / 1e9 / t_hot if t_hot else None,
                "unit": "Gbp/s", "method": "gprof self time of the path's functions (saveInterval ... callPeaks) in a -pg build of "
                "the same source, scaled by wall / wall_pg; libc time (strtok, memset of calloc) is not sampled by gprof",
                "gprof_self_seconds_hot": round(hot, 3), "gprof_self_seconds_all": round(tot, 3), "wall_pg_s": round(wall_pg, 2)}

    def records(self):
        """the same fragments as interval records, per replicate (treatment, control or None)"""
        from genrich_b200 import host
        out = []
        for t, c in self.wls:
            out.append((host.fragments_to_intervals(t.fragments(), atac=self.wl["atac"]),
                        host.fragments_to_intervals(c.fragments(), atac=self.wl["atac"]) if c is not None else None))
        return out

    def gpu_check(self, api, capi, host):
        """push the sample through the CUDA library and compare with the narrowPeak the reference wrote"""
        par = capi.make_params(p=self.wl["p"], q=self.wl["q"])
        ctx = capi.Context(api, self.spec["chrom_len"], par)
        res = host.run_replicates(ctx, self.records(), packed=True)
        ref = parse_narrowpeak(self.out)
        pk = res.peaks
        same = len(ref) == len(pk)
        worst_p = worst_q = 0.0
        if same and len(pk):
            r = np.array([x[:4] for x in ref], dtype=np.int64)
            g = np.stack([pk["chrom"].astype(np.int64), pk["start"], pk["end"], pk["summit"].astype(np.int64)], axis=1)
            same = bool(np.array_equal(r, g))
            if same:
                worst_p = float(np.max(np.abs(np.array([x[4] for x in ref]) - pk["pval"].astype(np.float64))))
                if self.wl["q"]:
                    worst_q = float(np.max(np.abs(np.array([x[5] for x in ref]) - pk["qval"].astype(np.float64))))
        ctx.close()
        return {"peaks_reference": len(ref), "peaks_gpu": int(len(pk)), "coords_identical": bool(same),
                "worst_dp": round(worst_p, 7), "worst_dq": round(worst_q, 7),
                "ok": bool(same and worst_p <= 1e-4 + 1e-6 and worst_q <= 1e-4 + 1e-6),   # + half a unit of the 6 printed decimals
                "args": " ".join(ref_args(self.wl)), "genome_bp": self.G,
                "fragments": [list(x) for x in self.spec["reps"]]}

    def cli_e2e(self, threads):
        """the drop-in host program on the same SAM files: wall clock of the whole program, decode included"""
        cli = os.path.join(ROOT, "genrich_b200", "bin", "genrich-b200")
        if not os.path.exists(cli):
            return None
        out = os.path.join(self.dir, "cli.narrowPeak")
        best, runs = None, []
        for _ in range(2):                            # the second run has the library and the files warm
            t0 = time.perf_counter()
            r = subprocess.run(self.cmd(cli, out) + ["--threads", str(threads)], stderr=subprocess.PIPE, text=True)
            dt = time.perf_counter() - t0
            if r.returncode:
                return {"error": (r.stderr or "")[-300:]}
            best = dt if best is None else min(best, dt)
            runs.append(round(dt, 3))
        ident = [l.split("\t")[:3] + l.rstrip("\n").split("\t")[9:] for l in open(out)] == \
                [l.split("\t")[:3] + l.rstrip("\n").split("\t")[9:] for l in open(self.out)]
        return {"seconds": round(best, 3), "runs_s": runs, "value": self.G / 1e9 / best, "unit": "Gbp/s", "threads": threads,
                "narrowpeak_cols_1_3_10_identical_to_reference": bool(ident),
                "note": "genrich-b200 (host C program over the CUDA library) on the SAM files the reference was timed on: "
                        "whole program, SAM decode on the host threads included"}

    def cleanup(self):
        for f in os.listdir(self.dir):
            os.unlink(os.path.join(self.dir, f))
        os.rmdir(self.dir)


def describe_sample(spec, sec, npk):
    return "%d chrom x %d bp, fragments per replicate (treatment, control) %s, same generator and depth as the workload; " \
           "whole program incl. SAM parse, %.2f s/run, %d peaks" % (
               len(spec["chrom_len"]), spec["chrom_len"][0], [list(x) for x in spec["reps"]], sec, npk)


def reference_arm(a, wl):
    rs = RefSample(wl["sample"], wl, a.workload)
    try:
        runs = max(1, min(a.steps, 8))                 # ~13 s each: the whole arm stays within a few minutes whatever --steps says
        times = rs.run_reference(1 + runs)[1:]
        sec = sum(times) / len(times)
        npk = sum(1 for _ in open(rs.out))
        cb = {"value": rs.G / 1e9 / sec, "unit": "Gbp/s", "cores": 1, "kind": "reference",
              "sample": describe_sample(wl["sample"], sec, npk), "host_cores_available": os.cpu_count(),
              "runs_timed": len(times), "hot_path": rs.hot_path(sec)}
    finally:
        rs.cleanup()
    return cb, sec


# --------------------------------------------------------------------------------------
def load_dram_table(workload, world):
    p = os.path.join(ROOT, "profiles", "r02_dram_by_stage.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t.get("%s@%d" % (workload, world))


def main():
    import faulthandler
    faulthandler.enable()                            # a crash inside a library says where it was called from
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="hg38_chip_50M_50M")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference run, the parity gate and the host-program timing")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-formulation (GR_FUSED=0) and CTA-scan comparison passes")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="warm-up + ONE step and nothing else (for ncu); writes gpurun_out/profile_meta.json")
    ap.add_argument("--prefetch-depth", type=int, default=2, help="e2e arm: samples sent ahead of their push (<= 2)")
    ap.add_argument("--gen-threads", type=int, default=0)
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
        cb, sec = reference_arm(a, wl)
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

    numa = bind_to_gpu_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                # comes up while the workload is generated
    host_group = None
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
        host_group = td.new_group(backend="gloo")      # host-resident scalars
    api = capi.load_cuda()
    par = capi.make_params(p=wl["p"], q=wl["q"])
    threads = a.gen_threads or max(2, min(16, (os.cpu_count() or 8) // max(world, 1)))

    # N > 1 first proves the sharded path on a small -q workload with peaks: every rank computes the whole genome
    # alone, then the ranks do it together (chromosomes sharded, sums all-reduced, BH histogram all-gathered,
    # peaks gathered) -- same records, so rank 0 must see the same peaks, bit for bit.  16 chromosomes: every rank
    # of an 8-rank job owns some (a context that owns nothing is refused by gr_create).
    shard_parity = None
    if world > 1 and not a.profile:
        mw = dict(chrom_len=[6_000_000] * 16, reps=[(1_500_000, 1_500_000)], enrich=0.4, spacing=200000, sigma=60.0)
        mpar = capi.make_params(p=None, q=0.05)
        from genrich_b200.synth import Workload
        mt = Workload(mw["chrom_len"], mw["reps"][0][0], 4001, enrich=mw["enrich"], spacing=mw["spacing"], sigma=mw["sigma"]).fragments()
        mc = Workload(mw["chrom_len"], mw["reps"][0][1], 4002, enrich=0.0).fragments()
        single_ctx = capi.Context(api, mw["chrom_len"], mpar, device=local)       # a world-1 context inside a world-N job
        want = host.run_replicates(single_ctx, [(mt, mc)], packed=True).peaks.copy()
        single_ctx.close()
        meng = ShardedEngine(api, mw["chrom_len"], mpar, dev, host_group=host_group)
        mt_r, mc_r = meng.route(mt), meng.route(mc)
        pt, _ = host.pack_records(mt_r)
        pc, _ = host.pack_records(mc_r)
        meng.replicate(lambda c: c.push_packed(pt), lambda c: c.push_packed(pc), want_stats=False)
        got, mrs = meng.call_peaks()
        if rank == 0:
            shard_parity = {"workload": "16 x 6 Mbp, 1.5 M + 1.5 M fragments, -q 0.05 (BH histogram all-gather on the path)", "peaks": int(len(got)),
                            "identical_to_single_context": bool(got.tobytes() == want.tobytes()),
                            "distinct_p": int(mrs.n_distinct_p), "hist_allgather_bytes": int(meng.hist_bytes)}
            assert shard_parity["identical_to_single_context"] and len(got) > 0, "sharded run differs from the single-context run"
        meng.close()
        del meng

    eng = ShardedEngine(api, L, par, dev, host_group=host_group)
    ctx = eng.ctx
    use6 = ctx.pack6_layout() is not None

    # synthetic interval records of this rank's chromosomes: 8-byte words in HBM, 6-byte records in pinned host memory
    t_gen0 = time.perf_counter()
    reps = []
    for r, (nt, nc) in enumerate(wl["reps"]):
        tag = "%s_w%d_r%d_rep%d" % (a.workload, world, rank, r)
        t = Sample(torch, host, eng, wl, nt, 2001 + 10 * r, wl["enrich"], dev, use6, threads, tag + "_t")
        c = Sample(torch, host, eng, wl, nc, 2002 + 10 * r, 0.0, dev, use6, threads, tag + "_c") if nc else None
        reps.append((t, c))
    order = [s for rp in reps for s in rp if s is not None]
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t_gen0
    feeder = Feeder(order, max(0, min(2, a.prefetch_depth)))

    def step(from_host, eng=eng):
        eng.ctx.reset()
        eng.saved_any[:] = False
        eng.sample_stats.clear()
        for t, c in reps:
            if from_host:
                eng.replicate(feeder.push, feeder.push if c is not None else None, want_stats=False)
            else:
                eng.replicate(t.push_dev, c.push_dev if c is not None else None, want_stats=False)
        if from_host:
            feeder.top_up(eng.ctx)                     # the next step's first samples travel under the peak calling
        return eng.call_peaks()

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
        # the engine hands out buffers it reuses (valid until its next call): keep a copy, outside the timed region
        return float(t[0]), float(t[1]), ctx.kernel_launches() - l0, peaks.copy(), rs, stages

    if a.profile:
        # for ncu: the launches of the LAST step are what tools/ncu_dram_by_stage.py keeps
        _, _, launches, peaks, rs, _ = timed(False, 1, 3)
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump({"workload": a.workload, "world": world, "launches_per_step": int(launches), "peaks": int(len(peaks))},
                      open(os.path.join(ROOT, "gpurun_out", "profile_meta.json"), "w"))
        sampler.stop()
        if world > 1:
            td.destroy_process_group()
        return

    t_w0 = time.time()
    # headline: K steps, nothing but the step itself in the stream; the per-stage CUDA events (two
    # per stage, ~20 stages per step) are recorded in a separate short pass of the same step
    ms_dev, wall_dev, launches, peaks, rs, _ = timed(False, a.steps, a.warmup)
    host_phase_dev = {k: round(v * 1e3 / a.steps, 4) for k, v in eng.t_acc.items()}
    ms_e2e = wall_e2e = None
    peaks2 = peaks
    host_phase = {}
    if not a.no_e2e:
        ms_e2e, wall_e2e, _, peaks2, _, _ = timed(True, a.steps, 2)
        # rank 0's host time by phase of the e2e arm (where the host waits for the device it is device time too)
        host_phase = {k: round(v * 1e3 / a.steps, 4) for k, v in eng.t_acc.items()}
    st_steps = max(2, a.steps // 2)
    ms_staged, _, _, peaks3, _, stages = timed(False, st_steps, 1, with_stages=True)
    assert peaks3.tobytes() == peaks.tobytes()
    clocks = sampler.stop(t_w0, time.time()) if rank == 0 else None
    n_breaks = int(sum(ctx.replicate_stats(r).n_pval for r in range(len(reps)))) if rank == 0 else 0
    hist_bytes = eng.hist_bytes

    # the same per-base pass in its two other formulations, same records, peaks asserted identical:
    #   dense  (GR_FUSED=0): the delta array is written to HBM by k_sb_build and read back by k_scan_stream -- the
    #          kernel the HBM-read roofline is literally about
    #   cta    (GR_FUSED_CTA=1): k_fb_scan, the CTA-owned shared-memory cell array that was the default in round 1
    forms = {}
    if world == 1 and not a.no_dense and G < (1 << 32):
        for key, env in (("dense_array", {"GR_FUSED": "0"}), ("cta", {"GR_FUSED_CTA": "1"}), ("rank", {"GR_FUSED_RANK": "1"})):
            os.environ.update(env)
            eng_d = ShardedEngine(api, L, par, dev, host_group=None)
            ms_d, _, _, peaks_d, _, st_d = timed(False, st_steps, 3, with_stages=True, eng=eng_d)
            for k in env:
                del os.environ[k]
            assert peaks_d.tobytes() == peaks.tobytes(), "%s and default formulations disagree" % key
            forms[key] = (ms_d, st_d)
            eng_d.close()
            del eng_d
    dense, cta = forms.get("dense_array"), forms.get("cta")

    if eng.debug:
        print("rank %d host-side ms per step (e2e arm): %s" % (rank, host_phase), file=sys.stderr, flush=True)
    if rank != 0:
        eng.close()
        if world > 1:
            td.destroy_process_group()
        return
    n_samples = len(order)
    n_records = int(sum(s.n for s in order))
    peak_gbs, peak_src = measured_peak_gbs()
    fused = "fused_scan" in stages
    cells = sum((int(l) + 1 + 8191) // 8192 * 8192 for l, o in zip(L, eng.owned) if o)
    # the scan form the LAST sample of the step chose on the device (both kernels are launched, one returns at once)
    form, hot_entries, all_entries = ctx.scan_form()
    scan_kernel = ("k_fb_scan" if form == 1 else "k_fr_scan") if fused else "k_scan_stream"
    scan_ms, scan_launches, _ = stages.get("fused_scan" if fused else "dense_scan", (0.0, 0, 0))
    per_launch_ms = scan_ms / max(scan_launches, 1)          # mean over every launch of the staged pass
    place_ms, place_launches, _ = stages.get("scan_place", (0.0, 0, 0))
    achieved = 4.0 * cells / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms else 0.0

    def formulation(kernel, res, stage, note):
        if res is None:
            return None
        d_ms, d_n, _ = res[1].get(stage, (0.0, 0, 0))
        b_ms, b_n, _ = res[1].get("build", (0.0, 0, 0))
        d_per = d_ms / max(d_n, 1)
        d_ach = 4.0 * cells / (d_per * 1e-3) / 1e9 if d_per else 0.0
        return {"kernel": kernel, "achieved": d_ach, "peak": peak_gbs, "unit": "GB/s", "frac": d_ach / peak_gbs if peak_gbs else None,
                "ms_per_launch": d_per, "build_ms_per_launch": b_ms / max(b_n, 1) if b_n else None, "ms_per_step": res[0], "note": note}
    dense_obj = formulation("k_scan_stream", dense, "dense_scan",
                            "GR_FUSED=0: the delta array is written to HBM by k_sb_build and read back by k_scan_stream "
                            "(4 B per cell each way: here `achieved` IS the HBM read rate); same peaks, bit for bit")
    cta_obj = formulation("k_fb_scan", cta, "fused_scan", "GR_FUSED_CTA=1: round 1's default scan; same peaks, bit for bit")
    rank_obj = formulation("k_fr_scan", forms.get("rank"), "fused_scan", "GR_FUSED_RANK=1: the rank form whatever the blocks hold")

    # what really moves through DRAM: per kernel, from the committed ncu pass of `bench.py --profile` on this workload
    table = load_dram_table(a.workload, world)
    traffic = frac_dram = None
    # what the per-base pass cannot avoid moving: the records in, its breaks and one bit per cell out
    info = 8 * n_records + 8 * n_breaks + (cells // 8) * n_samples
    step_obj = {"info_lower_bound_bytes": int(info),
                "info_note": "8 B per record in + 8 B per break out (breaks ~ the p intervals of every replicate) + 1 bit per cell and sample",
                "frac_info": info / (ms_dev * 1e-3) / 1e9 / peak_gbs if peak_gbs else None}
    if table is not None:
        ks = table["kernels"]
        # the kernel that did not run is in the table too (it returns at once: a few KB); a table taken before a
        # change of the chosen form has no useful entry: traffic stays null rather than quote another kernel's bytes
        k = ks.get(scan_kernel)
        if k is not None and k["launches"]:
            traffic = (k["dram_read"] + k["dram_write"]) / max(k["launches"], 1)
            frac_dram = traffic / (per_launch_ms * 1e-3) / 1e9 / peak_gbs if per_launch_ms else None
        tot = sum(v["dram_read"] + v["dram_write"] for v in ks.values())
        step_obj.update({"dram_bytes": int(tot), "frac_dram": tot / (ms_dev * 1e-3) / 1e9 / peak_gbs,
                         "dram_bytes_by_kernel": {n: int(v["dram_read"] + v["dram_write"]) for n, v in
                                                  sorted(ks.items(), key=lambda kv: -(kv[1]["dram_read"] + kv[1]["dram_write"]))[:12]},
                         "source": "profiles/r02_dram_by_stage.json (%s): dram__bytes_read.sum + dram__bytes_write.sum of every "
                                   "kernel of one step under ncu; memsets and copies are not kernels and are not in it" % table.get("command", "")})
    stage_ms = {k: round(v[0] / st_steps, 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])}
    stage_ms["_step_with_stage_events"] = round(ms_staged, 4)
    thr = "-q %g" % wl["q"] if wl["q"] else "-p %g" % wl["p"]
    line = {
        "metric": "Gbp p-value-scanned/sec", "value": G / 1e9 / (ms_dev * 1e-3), "unit": "Gbp/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 deltas, f32 pileups, f64 -> f32 -log10 p", "data": "synthetic",
        "config": {"workload": a.workload, "genome_bp": G, "chromosomes": len(L),
                   "fragments_per_replicate": [list(x) for x in wl["reps"]], "records_rank0": n_records,
                   "threshold": thr, "atac": wl["atac"], "multimap_fraction": wl["multimap"],
                   "sharding": "chromosomes over %d rank(s), LPT" % world, "numa_binding_rank0": numa,
                   "l2": "inputs (%.1f GB dense delta array per sample, %.2f GB of records) far exceed the 126 MB L2" % (
                       4e-9 * cells, 8e-9 * n_records),
                   "peaks": int(len(peaks)), "peaks_sha256": peaks_sha256(peaks), "intervals_rank0": int(rs.n_intervals),
                   "distinct_p": int(rs.n_distinct_p), "generation_s": round(gen_s, 1)},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall_dev,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs if peak_gbs else None, "traffic": traffic, "frac_dram": frac_dram,
                     "peak_source": peak_src, "bytes_per_launch": 4 * cells, "ms_per_launch": per_launch_ms,
                     "companion_scan_place_ms_per_launch": place_ms / max(place_launches, 1),
                     "launches_per_step": scan_launches / st_steps, "samples_scanned_per_step": n_samples,
                     "scan_form": {"chosen": scan_kernel, "entries_in_full_blocks": int(hot_entries), "entries": int(all_entries),
                                   "rule": "chosen on the device per sample: CTA form when a quarter of the entries lie in blocks of "
                                           ">= 1024 entries, else rank form; both kernels are launched, the other returns at once "
                                           "(its launch is inside ms_per_launch)"},
                     "note": "`achieved` / `frac` divide the ALGORITHMIC bytes of the per-base pass (SURVEY 8d: 4 B per base per "
                             "sample array) by the launch time.  The delta cells of the fused scan never exist in HBM (bucketed events "
                             "in, breaks and a bitmap out), so frac > 1 is not a bandwidth claim: `traffic` / `frac_dram` say what "
                             "the kernel really moves, `step` what the whole step moves, `dense_formulation` what the kernel that "
                             "does read 4 B per cell achieves",
                     "step": step_obj, "dense_formulation": dense_obj, "cta_formulation": cta_obj,
                     "rank_formulation": rank_obj},
        "dense_formulation": dense_obj,
        "stage_ms_per_step": stage_ms,
        "host_phase_ms_per_step_device_arm": host_phase_dev,
        "host_phase_ms_per_step_e2e": host_phase,
        "shard_parity": shard_parity,
    }
    if not a.no_e2e:
        h2d = int(sum(s.host_bytes for s in order))
        line["e2e"] = {"value": G / 1e9 / (ms_e2e * 1e-3), "unit": "Gbp/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(peaks2.nbytes + 512),
                       "record_format": "GR_PACK6 (6 B per record, expanded on the device)" if order[0].fmt == 6 else "GR_PACK (8 B per record)",
                       "wall_ms_per_step": wall_e2e, "prefetch_depth": feeder.depth,
                       "peaks_identical_to_device_arm": bool(peaks2.tobytes() == peaks.tobytes())}
    if wl["q"]:
        line["bh"] = {"distinct_p": int(rs.n_distinct_p), "hist_allgather_bytes_per_step": int(hist_bytes),
                      "stage_ms": {k: stage_ms.get(k) for k in ("bh_hist", "bh")},
                      "host_phase_ms": {k: host_phase_dev.get(k) for k in ("bh_sizes", "bh_allgather", "bh_exchange_and_q")}}
    eng.close()                                    # torch objects first, then the context and its stream
    if world > 1:
        td.destroy_process_group()                 # nothing below involves the other ranks
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Genrich"))
    if world == 1 and not a.no_cpu_baseline and have_ref:
        # the contract: on rank 0 at N = 1 only.  One reference run per sample; the workload's own sample is the
        # CPU baseline, and every sample goes through the GPU for the parity gate.
        rsm = RefSample(wl["sample"], wl, a.workload)
        try:
            sec = rsm.run_reference(1)[0]
            npk = sum(1 for _ in open(rsm.out))
            line["cpu_baseline"] = {"value": rsm.G / 1e9 / sec, "unit": "Gbp/s", "cores": 1, "kind": "reference",
                                    "sample": describe_sample(wl["sample"], sec, npk),
                                    "host_cores_available": os.cpu_count(), "hot_path": rsm.hot_path(sec)}
            gate = {a.workload + " sample": rsm.gpu_check(api, capi, host)}
            line["cli_e2e"] = rsm.cli_e2e(min(16, os.cpu_count() or 8))
            if line["cli_e2e"] and "seconds" in line["cli_e2e"]:
                line["cli_e2e"]["reference_seconds"] = round(sec, 3)
                line["cli_e2e"]["speedup_vs_reference"] = round(sec / line["cli_e2e"]["seconds"], 2)
        finally:
            rsm.cleanup()
        for name, spec in (GATE_EXTRA.items() if a.workload == "hg38_chip_50M_50M" else ()):
            g = RefSample(dict(chrom_len=spec["chrom_len"], reps=spec["reps"]), spec, name)
            try:
                g.run_reference(1)
                gate[name] = g.gpu_check(api, capi, host)
            finally:
                g.cleanup()
        line["parity"] = {"identical": all(v["ok"] for v in gate.values()),
                          "peaks": sum(v["peaks_gpu"] for v in gate.values()),
                          "worst_dp": max(v["worst_dp"] for v in gate.values()),
                          "worst_dq": max(v["worst_dq"] for v in gate.values()),
                          "rule": "narrowPeak of the unmodified reference, written in this run on the SAM view of each sample, vs the "
                                  "CUDA path on the same fragments: line count, columns 1-3 and 10 identical, columns 8-9 within 1e-4",
                          "samples": gate}
    else:
        line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": 1, "kind": "reference",
                                "sample": "skipped (N > 1, --no-cpu-baseline or oracle/_ref missing)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
