"""Multi-GPU dispatcher: chromosomes sharded over ranks, one process per GPU.

This is what runProgram's per-chromosome loop (Genrich.c:5460-5607) becomes.  Every
rank owns a subset of the chromosomes (greedy LPT by length) and runs the whole
hot path on them; only three things cross ranks, all through torch.distributed:

  * per-chromosome sum(len*val) of the experimental pileup  -> lambda   (all_reduce, nchrom doubles)
  * per-chromosome sum(len*val) of the control pileup       -> factor   (all_reduce, nchrom doubles)
  * the histogram of distinct -log10 p that Benjamini-Hochberg needs    (all_gather, -q only)

Each chromosome has exactly one non-zero contributor, so the all_reduce is exact,
and the totals are then added in chromosome order on every rank: lambda and the
scale factor are bit-identical to the single-GPU run.  Peaks come back to rank 0
in chromosome order (callPeaks numbers them globally, Genrich.c:986).

The engine is whatever :class:`~genrich_b200.capi.Api` the caller passes -- the CUDA
library in production (NCCL backend), and in the CPU test-suite the oracle twin
(gloo backend), which exercises exactly this host logic.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as td

from .capi import Api, Context, GrParams, PEAK_DTYPE
from .host import lpt_shard


class _CudaView:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def _tensor_from_ptr(ptr: int, n: int, np_dtype, device: torch.device) -> torch.Tensor:
    if n == 0:
        return torch.empty(0, dtype=torch.from_numpy(np.empty(0, np_dtype)).dtype, device=device)
    if device.type == "cuda":
        return torch.as_tensor(_CudaView(ptr, n, np.dtype(np_dtype).str), device=device)
    buf = (C.c_char * (n * np.dtype(np_dtype).itemsize)).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np_dtype, count=n).copy())


class ShardedEngine:
    def __init__(self, api: Api, chrom_len, params: GrParams, device: torch.device, skip=None, host_group=None,
                 exclusions=None):
        """host_group: process group for the small HOST-side exchanges (two per-chromosome
        double vectors per replicate, the peak records).  With NCCL as the default backend
        pass a gloo group (``td.new_group(backend="gloo")``): these values live on the host,
        and a loopback exchange costs ~0.1 ms where an NCCL call costs a device round trip
        plus two stream syncs.  The BH histogram -- the one device-resident exchange --
        always goes through the default (NCCL) group.
        exclusions: -E regions as (chromosome index, start, end) records; every rank gets them all."""
        self.world = td.get_world_size() if td.is_initialized() else 1
        self.rank = td.get_rank() if td.is_initialized() else 0
        self.device = device
        self.host_group = host_group
        self.chrom_len = np.asarray(chrom_len, dtype=np.uint32)
        self.nchrom = len(chrom_len)
        self.skip = np.zeros(self.nchrom, np.uint8) if skip is None else np.asarray(skip, np.uint8)
        self.owner = lpt_shard(np.where(self.skip != 0, 0, self.chrom_len), self.world)
        self.owned = (self.owner == self.rank).astype(np.uint8)
        dev_index = device.index if device.type == "cuda" and device.index is not None else 0
        if api.has_device and not np.any((self.owned != 0) & (self.skip == 0) & (self.chrom_len > 0)):
            raise ValueError("rank %d of %d would own no chromosome (%d in the table): use at most as many ranks as "
                             "chromosomes -- gr_create refuses a context without analyzable genome" %
                             (self.rank, self.world, self.nchrom))
        self.ctx = Context(api, chrom_len, params, device=dev_index, skip=self.skip, owned=self.owned)
        self.params = params
        self.excluded = np.zeros(self.nchrom, dtype=np.int64)      # bp per chromosome (saveXBed-merged)
        if exclusions is not None and len(exclusions):
            self.ctx.set_exclusions(exclusions)
            self.excluded = self.ctx.excluded_bp().astype(np.int64)
        self.saved_any = np.zeros(self.nchrom, dtype=bool)
        self.sample_stats = []
        self._ext_stream = None
        self._dsums = None
        self._gather_cap = 1 << 13          # peak records per rank in the gather slot (grown on demand)
        self._send = self._recv = self._host = self._merged = None
        self._slot_bytes = 0
        self._stage = None
        self.debug = bool(os.environ.get("GR_DIST_DEBUG"))
        self.t_acc = {}
        self.hist_bytes = 0                 # bytes the last BH histogram all-gather moved (all ranks)
        self._keep = None
        self._hist_cap = 0
        self._hist_buf = None

    def _tick(self, name, t0):
        # host time by phase, always collected (two clock reads per phase); bench.py reports it
        self.t_acc[name] = self.t_acc.get(name, 0.0) + (time.perf_counter() - t0)

    # records of chromosomes this rank does not own are dropped here (host routing)
    def route(self, recs: np.ndarray) -> np.ndarray:
        recs = np.ascontiguousarray(recs, dtype=np.int32).reshape(-1, 4)
        if self.world == 1:
            return recs
        return recs[self.owned[recs[:, 0]] != 0]

    def _reduce_sums(self, sums: np.ndarray) -> float:
        t0 = time.perf_counter()
        if self.world > 1:
            if self.host_group is not None or self.device.type == "cpu":
                t = torch.from_numpy(np.ascontiguousarray(sums))
                td.all_reduce(t, op=td.ReduceOp.SUM, group=self.host_group)
                sums = t.numpy()
            else:
                t = torch.from_numpy(sums).to(self.device)
                td.all_reduce(t, op=td.ReduceOp.SUM)
                sums = t.cpu().numpy()
        tot = 0.0
        for v in sums:                       # chromosome order, like the reference's running sum
            tot += float(v)
        self._tick("reduce_sums", t0)
        return tot

    def replicate(self, push_expt, push_ctrl=None, save=None, want_stats=True):
        """push_*: callables that feed this rank's records into self.ctx.

        CUDA library with want_stats=False: nothing waits for the device -- both pileups are
        enqueued, the per-chromosome sums are all-reduced where they are (NCCL on the library's
        own stream) and lambda / the scale factor are computed on the device."""
        sv = np.ones(self.nchrom, np.uint8) if save is None else np.asarray(save, np.uint8)
        self.saved_any |= (sv != 0) & (self.skip == 0)
        act = (sv != 0) & (self.skip == 0)
        glen = int((self.chrom_len[act].astype(np.int64) - self.excluded[act]).sum())      # calcLambda 1819-1827
        if self.ctx.api.has_device and self.device.type == "cuda" and not want_stats:
            t0 = time.perf_counter()
            self.ctx.sample_begin(False, sv)
            push_expt(self.ctx)
            self.ctx.sample_pileup_async()
            if push_ctrl is not None:
                self.ctx.sample_begin(True)
                push_ctrl(self.ctx)
                self.ctx.sample_pileup_async()
            self._tick("push_pileup", t0)
            t0 = time.perf_counter()
            if self.world > 1:
                if self._ext_stream is None:
                    self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream_handle(), device=self.device)
                    ep, _ = self.ctx.sums_device_ptrs()
                    self._dsums = _tensor_from_ptr(ep, 2 * self.nchrom, np.float64, self.device)
                with torch.cuda.stream(self._ext_stream):
                    td.all_reduce(self._dsums, op=td.ReduceOp.SUM)      # every chromosome has one owner: exact
            self._tick("reduce_sums", t0)
            t0 = time.perf_counter()
            self.ctx.replicate_finish_device(push_ctrl is not None, glen)
            self._tick("replicate_finish", t0)
            self.sample_stats.append(None)
            return None
        t0 = time.perf_counter()
        self.ctx.sample_begin(False, sv)
        push_expt(self.ctx)
        s0 = self.ctx.sample_pileup()
        self._tick("expt_push_pileup", t0)
        frag = self._reduce_sums(s0)
        ctrl = 0.0
        if push_ctrl is not None:
            t0 = time.perf_counter()
            self.ctx.sample_begin(True)
            push_ctrl(self.ctx)
            s1 = self.ctx.sample_pileup()
            self._tick("ctrl_push_pileup", t0)
            ctrl = self._reduce_sums(s1)
        t0 = time.perf_counter()
        st = self.ctx.replicate_finish(frag, ctrl, push_ctrl is not None, glen)
        self._tick("replicate_finish", t0)
        self.sample_stats.append(st)
        return st

    def _exchange_histogram(self):
        """The one data-path collective: all-gather of every rank's (key, bp) histogram of distinct
        -log10 p (computeQval 352 sees one genome-wide table).

        gr_bh_local_hist returns with the list complete (it waits for its stream), and the host
        knows its length.  CUDA: the lengths travel through the host group (or one small device
        all-gather), then ONE NCCL all-gather of fixed-size slots [keys | lens] issued on the
        library's own stream, so that gr_bh_set_global, which runs on that stream, is ordered
        behind it without any host synchronisation."""
        kp, lp, n = self.ctx.bh_local_hist_ptrs()
        keys = _tensor_from_ptr(kp, n, np.uint32, self.device).view(torch.int32)
        lens = _tensor_from_ptr(lp, n, np.uint64, self.device).view(torch.int64)
        if self.world == 1:
            return keys, lens
        cuda = self.device.type == "cuda"
        t0 = time.perf_counter()
        if not cuda or (self.host_group is not None and os.environ.get("GR_DIST_HOST_SIZES")):
            # (on GPUs the eight-byte exchange through the gloo group measured 0.5 ms at 8 ranks: NCCL below)
            cnt = torch.tensor([n], dtype=torch.int64)
            cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
            td.all_gather(cnts, cnt, group=self.host_group)
            sizes = [int(c[0]) for c in cnts]
        else:
            cnt = torch.tensor([n], dtype=torch.int64, device=self.device)
            cnts = torch.empty(self.world, dtype=torch.int64, device=self.device)
            td.all_gather_into_tensor(cnts, cnt)
            sizes = [int(v) for v in cnts.tolist()]
        self._tick("bh_sizes", t0)
        t0 = time.perf_counter()
        if not cuda:
            m = max(max(sizes), 1)
            kpad = torch.zeros(m, dtype=torch.int32)
            lpad = torch.zeros(m, dtype=torch.int64)
            kpad[:n] = keys
            lpad[:n] = lens
            kall = [torch.empty_like(kpad) for _ in range(self.world)]
            lall = [torch.empty_like(lpad) for _ in range(self.world)]
            td.all_gather(kall, kpad, group=self.host_group)
            td.all_gather(lall, lpad, group=self.host_group)
            keys = torch.cat([k[:s] for k, s in zip(kall, sizes)]).contiguous()
            lens = torch.cat([l[:s] for l, s in zip(lall, sizes)]).contiguous()
            self._tick("bh_allgather", t0)
            return keys, lens
        need = (max(max(sizes), 1) + 1) & ~1                # even: the lens part of a slot stays 8-byte aligned
        tot = sum(sizes)
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream_handle(), device=self.device)
        # The buffers are allocated on torch's current stream (memory that the caching allocator ties to the library's
        # stream would outlive that stream when the context is closed), sized with room to spare and kept from step to
        # step: they are only ever written and read on the library's stream, so reuse needs no synchronisation, and
        # what lies behind a rank's own n entries in its slot is never looked at.
        if self._hist_cap < need:
            self._hist_cap = max(need, 2 * self._hist_cap)
            m = self._hist_cap
            self._hist_buf = (torch.zeros(12 * m, dtype=torch.uint8, device=self.device),
                              torch.empty(self.world * 12 * m, dtype=torch.uint8, device=self.device),
                              torch.empty(self.world * m, dtype=torch.int32, device=self.device),
                              torch.empty(self.world * m, dtype=torch.int64, device=self.device))
            torch.cuda.current_stream(self.device).synchronize()   # the fill is done before the other stream writes
        m = self._hist_cap
        slot = 12 * m
        send, recv, keys_all, lens_all = self._hist_buf
        with torch.cuda.stream(self._ext_stream):
            if n:
                send[:4 * n].view(torch.int32).copy_(keys)
                send[4 * m:4 * m + 8 * n].view(torch.int64).copy_(lens)
            td.all_gather_into_tensor(recv, send)            # NCCL over NVLink, on the library's stream
            rv = recv.view(self.world, slot)
            at = 0
            for r, sz in enumerate(sizes):
                if sz:
                    keys_all[at:at + sz].copy_(rv[r, :4 * sz].view(torch.int32))
                    lens_all[at:at + sz].copy_(rv[r, 4 * m:4 * m + 8 * sz].view(torch.int64))
                at += sz
        self.hist_bytes = self.world * slot
        self._tick("bh_allgather", t0)
        return keys_all[:tot], lens_all[:tot]

    def close(self):
        """Release what refers to the library's stream and buffers, THEN the context (its stream goes with it)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        self._keep = self._hist_buf = self._send = self._recv = self._host = self._dsums = self._stage = None
        self._ext_stream = None
        self._slot_bytes = 0
        self._hist_cap = 0
        self.ctx.close()

    def _gather_peaks_one_wait(self):
        """CUDA, several ranks: the peak scan is enqueued (gr_call_peaks_enqueue), every rank's slot -- a 64-byte
        header the device fills in (count, bp, condition bits) followed by its records -- is all-gathered over
        NVLink on the LIBRARY's stream straight out of the library's buffer, rank 0 copies the lot to pinned host
        memory (the others only the headers), and then the host waits: once per step.  Every rank reads every
        header, so "a buffer was too small somewhere, again" is decided identically everywhere.  Returns None when
        a p-value table overflowed on some rank of a -q run (the caller starts over: p, exchange, q)."""
        t0 = time.perf_counter()
        isz = PEAK_DTYPE.itemsize
        if self._ext_stream is None:
            self._ext_stream = torch.cuda.ExternalStream(self.ctx.stream_handle(), device=self.device)
        while True:
            dslot, cap = self.ctx.call_peaks_enqueue()
            g = self._gather_cap                                          # the same on every rank (it only grows with the global counts)
            nb = 64 + g * isz
            if self._slot_bytes != nb:
                self._slot_bytes = nb
                self._recv = torch.empty(self.world * nb, dtype=torch.uint8, device=self.device)
                self._host = torch.empty(self.world * nb, dtype=torch.uint8, pin_memory=True)
                self._merged = np.empty(self.world * g, PEAK_DTYPE)
                self._stage = None
            with torch.cuda.stream(self._ext_stream):
                if cap >= g:
                    src = _tensor_from_ptr(dslot, nb, np.uint8, self.device)
                else:                                                     # a tiny shard: the library's buffer is shorter than the slot
                    if self._stage is None:
                        self._stage = torch.zeros(nb, dtype=torch.uint8, device=self.device)
                    have = 64 + cap * isz
                    self._stage[:have].copy_(_tensor_from_ptr(dslot, have, np.uint8, self.device), non_blocking=True)
                    src = self._stage
                td.all_gather_into_tensor(self._recv, src)               # NCCL over NVLink, on the library's stream
                if self.rank == 0:
                    self._host.copy_(self._recv, non_blocking=True)
                else:
                    self._host.view(self.world, nb)[:, :64].copy_(self._recv.view(self.world, nb)[:, :64], non_blocking=True)
                self._ext_stream.synchronize()                            # the one wait of the step
            host = self._host.numpy().reshape(self.world, nb)
            hdr = host[:, :64].copy().view(np.int64)                      # n_peaks, peak_bp, flags | reserved, n_intervals, pad
            flags = (hdr[:, 2] & 0xffffffff).astype(np.int64)
            redo, rs = self.ctx.call_peaks_done(host[self.rank, :64].ctypes.data)
            if np.any(flags & 32):                                        # GR_DE_TABLE on some rank: its p-values are redone ...
                if self.params.qval_opt:
                    return None                                           # ... and with them the histogram exchange (every rank)
                continue                                                  # ... by its next gr_call_peaks_enqueue
            counts = [int(c) for c in hdr[:, 0]]
            if np.any(flags & 64) or max(counts) > g:                     # GR_DE_CAP somewhere, or more peaks than the gather slot
                if max(counts) > g:
                    self._gather_cap = 2 * max(counts)
                continue
            break
        self._tick("call_peaks_gather", t0)
        t0 = time.perf_counter()
        peaks = np.empty(0, PEAK_DTYPE)
        if self.rank == 0:
            lists = (C.c_void_p * self.world)(*[host[r, 64:].ctypes.data for r in range(self.world)])
            cnts = (C.c_uint64 * self.world)(*counts)
            rc = self.ctx.api.merge_peaks(lists, cnts, self.world, self._merged.ctypes.data_as(C.c_void_p))
            if rc:
                raise RuntimeError("gr_merge_peaks failed: %d" % rc)
            peaks = self._merged[:sum(counts)]
        self._tick("merge_peaks", t0)
        return peaks, rs

    def call_peaks(self):
        """Returns (peaks, run_stats) -- peaks of ALL chromosomes on rank 0; elsewhere the rank's own peaks
        (host engines) or none (CUDA: they stay on the device)."""
        self.ctx.pvalues_finalize()
        if self.params.qval_opt:
            G = int(self.params.genome_len) or int((self.chrom_len[self.saved_any].astype(np.int64)
                                                    - self.excluded[self.saved_any]).sum())     # findPeaks 1091-1101
            t0 = time.perf_counter()
            keys, lens = self._exchange_histogram()
            self._keep = (keys, lens)
            # CUDA: the gathered lists were produced on the library's own stream (or, at one rank, by
            # gr_bh_local_hist, which returns with its stream idle): gr_bh_set_global is ordered behind them
            self.ctx.bh_set_global_ptrs(keys.data_ptr(), lens.data_ptr(), keys.numel(), G)
            self._tick("bh_exchange_and_q", t0)
        t0 = time.perf_counter()
        cuda_gather = self.world > 1 and self.device.type == "cuda"
        if cuda_gather:
            got = self._gather_peaks_one_wait()
            if got is None:                      # -q and a p-value table overflowed somewhere: p, the exchange and q once more
                return self.call_peaks()
            return got
        peaks, rs = self.ctx.call_peaks() if self.ctx.api.has_device else self.ctx.call_peaks()
        self._tick("call_peaks", t0)
        t0 = time.perf_counter()
        if self.world > 1:
            isz = PEAK_DTYPE.itemsize
            if cuda_gather:
                raise AssertionError("unreachable: the CUDA ranks gather in _gather_peaks_one_wait")
            else:
                buf = torch.from_numpy(peaks.view(np.uint8).copy())
                cnt = torch.tensor([buf.numel()], dtype=torch.int64)
                cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
                td.all_gather(cnts, cnt, group=self.host_group)
                sizes = [int(c.item()) for c in cnts]
                m = max(max(sizes), 1)
                pad = torch.zeros(m, dtype=torch.uint8)
                pad[:buf.numel()] = buf
                allb = torch.empty(self.world * m, dtype=torch.uint8)
                td.all_gather(list(allb.view(self.world, m).unbind(0)), pad, group=self.host_group)
                host = allb.numpy().reshape(self.world, m)
                if self.rank == 0:
                    # every rank's list is in (chromosome, start) order and a chromosome has one owner:
                    # the global list is the owners' per-chromosome runs in chromosome order
                    edges = []
                    for r, sz in enumerate(sizes):
                        col = np.ascontiguousarray(host[r, :sz].view(PEAK_DTYPE)["chrom"])
                        edges.append(np.searchsorted(col, np.arange(self.nchrom + 1)) * isz)
                    out = np.empty(sum(sizes), np.uint8)
                    pos = 0
                    for c in range(self.nchrom):
                        r = int(self.owner[c])
                        lo, hi = int(edges[r][c]), int(edges[r][c + 1])
                        if hi > lo:
                            out[pos:pos + hi - lo] = host[r, lo:hi]
                            pos += hi - lo
                    peaks = out[:pos].view(PEAK_DTYPE)
        self._tick("gather_peaks", t0)
        return peaks, rs
