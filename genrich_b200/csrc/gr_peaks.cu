// gr_peaks.cu -- K8: the peak scan (callPeaks 977-1069, updatePeak 943,
// checkPeak 916, resetVars 932) restated for parallel execution.
//
// The reference walks the intervals of a chromosome once, carrying one open peak.
// Equivalent formulation: the significant intervals (v > threshold) are linked
// into candidate peaks; two consecutive significant intervals of a chromosome
// stay in the same candidate unless a SKIP interval lies between them or the
// non-significant stretch between them is longer than maxGap (the test at 1032
// fires on the last such interval, whose end is the next significant start).
// Each candidate is then walked sequentially by one thread, in interval order,
// so the float AUC (950) and the summit rules (956-969) see the reference's
// exact operation order.  Candidates are independent, so peaks run in parallel.
#include "gr_tile.cuh"
#include "gr_internal.h"

#define PK_SKIP (-1.0f)

// ---- events: indices of significant or SKIP intervals, in order -----------------
// 8192 intervals per look-back tile: each warp takes 1024 consecutive values in 32
// coalesced rounds, ranks them with one ballot per round (order preserved) and
// keeps the 32 ballots in one register per lane for the write pass.
#define PE_TILE 8192
#define PK_GRID (148 * 8)        // persistent CTAs: tiles are taken by ticket until the (device-side) count is exhausted
__global__ void __launch_bounds__(256)
k_peak_events(const float* __restrict__ v, const u64* __restrict__ n_dev, float thr, Lookback<1> lb,
              u32* __restrict__ ev_idx, u64* __restrict__ ev_count) {
  const u64 n = *n_dev;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (;;) {
    const u32 tile = take_ticket(lb.ticket);
    if ((u64)tile * PE_TILE >= n) break;
    const u64 wbase = (u64)tile * PE_TILE + (u64)w * 1024;
    u32 mine = 0, cnt = 0;
#pragma unroll 8
    for (int k = 0; k < 32; k++) {
      const u64 i = wbase + k * 32 + lane;
      bool f = false;
      if (i < n) {
        const float x = v[i];
        f = x > thr || x == PK_SKIP;
      }
      const u32 bal = __ballot_sync(GR_FULL, f);
      if (lane == k) mine = bal;
      cnt += __popc(bal);                        // warp-uniform
    }
    // rank of the warp's first event: block scan over the 8 warp counts + look-back
    u32 tot;
    u64 r = tile_exclusive_rank(lb, tile, lane == 31 ? cnt : 0u, tot);   // lane 31 carries the warp's count
    r = __shfl_sync(GR_FULL, r, 31);             // exclusive rank of lane 31 == rank of the warp's first event
    for (int k = 0; k < 32; k++) {
      const u32 bal = __shfl_sync(GR_FULL, mine, k);
      if (bal & (1u << lane)) ev_idx[r + __popc(bal & ((1u << lane) - 1))] = (u32)(wbase + k * 32 + lane);
      r += __popc(bal);
    }
    if ((u64)(tile + 1) * PE_TILE >= n && threadIdx.x == 255) *ev_count = r;
    __syncthreads();                             // shared scratch of the ticket / rank helpers is reused
  }
}

// n_upper bounds the interval count (sizes the status words); the count is read from *n_dev
void launch_peak_events(cudaStream_t s, const float* v, u64 n_upper, const u64* n_dev, float thr, const PeakWork& w) {
  cudaMemsetAsync(w.ev_count, 0, sizeof(u64), s);
  if (!n_upper) return;
  const u64 ntiles = (n_upper + PE_TILE - 1) / PE_TILE;
  cudaMemsetAsync(w.sc.st, 0, (size_t)ntiles * sizeof(u64), s);
  cudaMemsetAsync(w.sc.ticket, 0, sizeof(u32), s);
  Lookback<1> lb;
  lb.st[0] = w.sc.st; lb.ticket = w.sc.ticket;
  k_peak_events<<<(unsigned)(ntiles < PK_GRID ? ntiles : PK_GRID), 256, 0, s>>>(v, n_dev, thr, lb, w.ev_idx, w.ev_count);
  GR_NOTE_LAUNCH();
}

// ---- heads: events that open a candidate ------------------------------------------
// 8 consecutive events per thread (2048 per tile): one ticket and one look-back per 2048.
#define PK_HEAD_PER 8
#define PK_HEAD_TILE (256 * PK_HEAD_PER)
__global__ void __launch_bounds__(256)
k_peak_heads(const u32* __restrict__ pEnd, const float* __restrict__ v,
             const u64* __restrict__ chrom_start, int nchrom, int max_gap,
             const u32* __restrict__ ev_idx, const u64* __restrict__ ev_count, Lookback<1> lb,
             u32* __restrict__ head_idx, u64* __restrict__ head_count) {
  const u64 nev = *ev_count;
  for (;;) {
    const u32 tile = take_ticket(lb.ticket);
    if ((u64)tile * PK_HEAD_TILE >= nev) break;
    const u64 t0 = ((u64)tile * 256 + threadIdx.x) * PK_HEAD_PER;
    u32 mask = 0;
    u32 prev = 0;
    bool have_prev = false, prev_skip = false;
    if (t0 > 0 && t0 < nev) { prev = ev_idx[t0 - 1]; have_prev = true; prev_skip = v[prev] == PK_SKIP; }
#pragma unroll
    for (int i = 0; i < PK_HEAD_PER; i++) {
      const u64 t = t0 + i;
      if (t >= nev) break;
      const u32 idx = ev_idx[t];
      const bool skip = v[idx] == PK_SKIP;
      u32 head = 0;
      if (!skip) {
        head = 1;
        if (have_prev && !prev_skip) {
          const int c = chrom_of_index(chrom_start, nchrom, idx);
          if ((u64)prev >= chrom_start[c]) {                 // same chromosome
            if (prev + 1 == idx) head = 0;                   // adjacent: no interval in between
            else {
              const i64 gap = (i64)pEnd[idx - 1] - (i64)pEnd[prev];   // start[idx] - peakEnd
              if (!(gap > (i64)max_gap)) head = 0;           // 1032
            }
          }
        }
      }
      mask |= head << i;
      prev = idx; have_prev = true; prev_skip = skip;
    }
    u32 tot;
    u64 r = tile_exclusive_rank(lb, tile, (u32)__popc(mask), tot);
#pragma unroll
    for (int i = 0; i < PK_HEAD_PER; i++)
      if (mask & (1u << i)) head_idx[r++] = (u32)(t0 + i);
    // the thread holding the last event reports
    if (t0 < nev && t0 + PK_HEAD_PER >= nev) *head_count = r;
    __syncthreads();
  }
}

// ---- walk: one thread per candidate ------------------------------------------------
// The arithmetic of a candidate is a chain (float AUC in interval order, summit rules), its loads are not: PK_FETCH
// events are fetched at a time so that their latencies overlap (a long candidate is one thread's critical path,
// and on a small shard the whole kernel's).  Candidates of one warp run side by side -- peaks come in runs of long
// candidates, so handing the long ones of a warp to the whole warp one after the other (tried: 4x slower) loses.
#define PK_FETCH 8
__global__ void __launch_bounds__(128)
k_peak_walk(const u32* __restrict__ pEnd, const float* __restrict__ pval,
            const float* __restrict__ qval, const u64* __restrict__ chrom_start, int nchrom,
            float thr, int qopt, float min_auc, int min_len,
            const u32* __restrict__ ev_idx, const u64* __restrict__ ev_count,
            const u32* __restrict__ head_idx, const u64* __restrict__ head_count,
            PeakRec* __restrict__ cand, uint8_t* __restrict__ cand_ok, u64 hcap, int* __restrict__ err) {
  const u64 nh = *head_count;
  if (nh > hcap) {                                       // candidate buffers were sized optimistically
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, GR_DE_CAP);
    return;
  }
  const u64 nev = *ev_count;
  const float* __restrict__ v = qopt ? qval : pval;
  for (u64 h = (u64)blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += (u64)gridDim.x * blockDim.x) {
    const u64 t0 = head_idx[h];
    const u64 t1 = h + 1 < nh ? head_idx[h + 1] : nev;
    const u32 first = ev_idx[t0];
    const int c = chrom_of_index(chrom_start, nchrom, first);
    const u64 cs = chrom_start[c];

    float auc = 0.0f, sVal = -1.0f, sP = -1.0f, sQ = -1.0f;     // 1001-1006
    i64 pStart = -1, pEndv = -1;
    u32 sPos = 0, sLen = 0;
    bool open = true;
    for (u64 t = t0; t < t1 && open; t += PK_FETCH) {
      u32 idx[PK_FETCH], end[PK_FETCH], start[PK_FETCH];
      float xv[PK_FETCH];
#pragma unroll
      for (int k = 0; k < PK_FETCH; k++) idx[k] = t + k < t1 ? ev_idx[t + k] : first;
#pragma unroll
      for (int k = 0; k < PK_FETCH; k++) {
        xv[k] = v[idx[k]];
        end[k] = pEnd[idx[k]];
        start[k] = (u64)idx[k] == cs ? 0u : pEnd[idx[k] - 1];
      }
#pragma unroll
      for (int k = 0; k < PK_FETCH; k++) {
        if (t + k >= t1) break;
        const float x = xv[k];
        if (x == PK_SKIP) { open = false; break; }                 // 1031: SKIP closes the candidate
        const u32 len = end[k] - start[k];
        auc = __fadd_rn(auc, __fmul_rn(__uint2float_rn(len), __fsub_rn(x, thr)));   // 950, no FMA
        if (pStart == -1) pStart = start[k];
        pEndv = end[k];
        if (x > sVal) {                                             // 956-961
          sVal = x;
          sP = pval[idx[k]];
          sQ = qopt ? qval[idx[k]] : PK_SKIP;
          sPos = (u32)((end[k] + start[k]) / 2 - (u32)pStart);      // uint32 arithmetic, 960
          sLen = len;
        } else if (x == sVal && len > sLen) {                       // 962-968
          sPos = (u32)((end[k] + start[k]) / 2 - (u32)pStart);
          sLen = len;
        }
      }
    }
    PeakRec r;
    r.chrom = c; r.summit = sPos; r.start = pStart; r.end = pEndv;
    r.auc = auc; r.pval = sP; r.qval = sQ; r.reserved = 0.0f;
    cand[h] = r;
    cand_ok[h] = (pStart != -1 && auc >= min_auc && pEndv - pStart >= (i64)min_len) ? 1 : 0;   // 920
  }
}

__global__ void __launch_bounds__(256)
k_peak_compact(const PeakRec* __restrict__ cand, const uint8_t* __restrict__ ok,
               const u64* __restrict__ head_count, Lookback<1> lb, PeakRec* __restrict__ out,
               u64* __restrict__ out_count, u64* __restrict__ peak_bp, u64 hcap) {
  const u64 nh = *head_count;
  if (nh > hcap) return;                                 // flagged by the walk
  for (;;) {
    const u32 tile = take_ticket(lb.ticket);
    if ((u64)tile * 256 >= nh) break;
    const u64 h = (u64)tile * 256 + threadIdx.x;
    const u32 f = h < nh && ok[h];
    u32 tot;
    const u64 r = tile_exclusive_rank(lb, tile, f, tot);
    u64 bp = 0;
    if (f) {
      const PeakRec p = cand[h];
      out[r] = p;
      bp = (u64)(p.end - p.start);                               // 924
    }
    bp = warp_sum_u64(bp);
    if ((threadIdx.x & 31) == 0 && bp) atomicAdd(peak_bp, bp);
    if (h + 1 == nh) *out_count = r + f;
    __syncthreads();
  }
}

// The three follow-up kernels take their sizes (event / head counts) from device memory:
// the host does not wait between the events kernel and the peak records.
void launch_peak_chain(cudaStream_t s, const u32* pEnd, const float* pval, const float* qval,
                       const u64* chrom_start, int nchrom, float thr, int qopt, int max_gap,
                       float min_auc, int min_len, const PeakWork& w, u64 nev_upper, u64 hcap, int* err) {
  cudaMemsetAsync(w.head_count, 0, sizeof(u64), s);
  cudaMemsetAsync(w.out_count, 0, sizeof(u64), s);
  cudaMemsetAsync(w.peak_bp, 0, sizeof(u64), s);
  if (!nev_upper) return;
  const float* v = qopt ? qval : pval;
  const u64 nth = (nev_upper + PK_HEAD_TILE - 1) / PK_HEAD_TILE;
  Lookback<1> lb;
  lb.st[0] = w.sc.st; lb.ticket = w.sc.ticket;
  cudaMemsetAsync(w.sc.st, 0, (size_t)nth * sizeof(u64), s);
  cudaMemsetAsync(w.sc.ticket, 0, sizeof(u32), s);
  k_peak_heads<<<(unsigned)(nth < PK_GRID ? nth : PK_GRID), 256, 0, s>>>(pEnd, v, chrom_start, nchrom, max_gap, w.ev_idx,
                                                                      w.ev_count, lb, w.head_idx, w.head_count);
  GR_NOTE_LAUNCH();
  k_peak_walk<<<PK_GRID, 128, 0, s>>>(pEnd, pval, qval, chrom_start, nchrom, thr, qopt, min_auc, min_len, w.ev_idx,
                                      w.ev_count, w.head_idx, w.head_count, w.cand, w.cand_ok, hcap, err);
  GR_NOTE_LAUNCH();
  const u64 ntc = (hcap + 255) / 256;
  cudaMemsetAsync(w.sc.st, 0, (size_t)ntc * sizeof(u64), s);
  cudaMemsetAsync(w.sc.ticket, 0, sizeof(u32), s);
  k_peak_compact<<<(unsigned)(ntc < PK_GRID ? ntc : PK_GRID), 256, 0, s>>>(w.cand, w.cand_ok, w.head_count, lb, w.out,
                                                                        w.out_count, w.peak_bp, hcap);
  GR_NOTE_LAUNCH();
}
