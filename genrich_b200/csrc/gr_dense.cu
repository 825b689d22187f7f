// gr_dense.cu -- the per-base kernels: K1 delta scatter, K2 single-pass dense
// prefix sum with break compaction, K2b exact weighted-length reduction.
//
// Replaces saveInterval's diff-array writes (Genrich.c:2575-2583) and the
// sequential O(genome) loops of savePileupExpt (2239-2273) / calcFactor
// (2013-2038).  HBM-bound int32 work: no tensor cores.
#include "gr_common.cuh"
#include "gr_internal.h"

// ============================================================================
// K1: two int32 reductions (RED.ADD) per interval record into the dense delta
// array, in units of 1/120 (weights 120/count, count in {1,2,3,4,5,6,8,10}:
// addFrac 2311 / subFrac 2412).  Clamping as saveInterval 2522-2544.
__global__ void __launch_bounds__(256)
k_scatter(const int4* __restrict__ recs, u64 n, DevLayout L, int32_t* __restrict__ delta,
          int* __restrict__ err, u64* __restrict__ clamped) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int4 r = ld_stream_v4(recs + i);
    const int c = r.x;
    if (c < 0 || c >= L.nchrom) { e_local |= GR_DE_CHROM; continue; }
    const uint8_t f = L.flags[c];
    const u64 off = L.off[c];
    if (off == ~0ull) {                      // not owned by this context, or -e skipped
      if (!(f & GR_CF_OWNED)) e_local |= GR_DE_CHROM;
      continue;
    }
    if (!(f & GR_CF_SAVE)) continue;         // processPair 3137-3138: not in this replicate
    const int cnt = r.w;
    if (cnt < 1 || cnt > 10 || !((1 << cnt) & 0x57E)) { e_local |= GR_DE_COUNT; continue; }
    const i64 len = L.len[c];
    i64 s = r.y, e = r.z;
    bool cl = false;
    if (s < 0) { s = 0; cl = true; }
    if (s >= len || e < 0) { e_local |= GR_DE_POS; continue; }
    if (e > len) { e = len; cl = true; }
    c_local += cl;
    const int w = 120 / cnt;
    atomicAdd(delta + off + s, w);
    atomicAdd(delta + off + e, -w);
  }
  if (e_local) atomicOr(err, e_local);
  if (c_local) atomicAdd(clamped, (u64)c_local);
}

void launch_scatter(cudaStream_t s, const DevLayout& L, const int32_t* recs, u64 n,
                    int32_t* delta, int* err, u64* clamped) {
  if (!n) return;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_scatter<<<(unsigned)blocks, 256, 0, s>>>((const int4*)recs, n, L, delta, err, clamped); GR_NOTE_LAUNCH();
}

// ============================================================================
// K2: one pass over the dense int32 delta cells.  Per 4096-cell tile:
//   coalesced LDG.128 (striped) -> per-warp padded shared-memory transpose ->
//   16 consecutive cells per thread -> thread/warp/block scan of (sum, #breaks)
//   -> decoupled look-back for the exclusive prefix of both -> breaks written
//   as (end, value) at their global rank + 1 bit per cell into the break bitmap.
// A break closes an interval at chromosome position j iff 1 <= j < len and
// delta[j] != 0, or j == len (Genrich.c:2241, 2268); its value is the running sum
// BEFORE delta[j] is added (2245), rebuilt as the reference float.
// Running sums are kept modulo 2^32: every true prefix fits in int32.
#define SCAN_THREADS 256
#define SCAN_WARPS 8
#define SCAN_ITEMS 16
#define SCAN_PAD_INT4 160      // 128 int4 per warp + 1 pad per 4

__global__ void __launch_bounds__(SCAN_THREADS)
k_dense_scan(const int32_t* __restrict__ delta, DevLayout L, Lookback<2> lb, DevRle out,
             u32* __restrict__ bitmap, int* __restrict__ err, u32 ntiles) {
  __shared__ int4 sm_x[SCAN_WARPS * SCAN_PAD_INT4];
  __shared__ u32 sm_wsum[SCAN_WARPS];
  __shared__ u32 sm_wcnt[SCAN_WARPS];
  __shared__ u32 sm_excl_sum;
  __shared__ u64 sm_excl_cnt;
  __shared__ u32 sm_tile;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) sm_tile = atomicAdd(lb.ticket, 1u);
  __syncthreads();
  const u32 tile = sm_tile;
  const u64 tbase = (u64)tile * GR_SCAN_TILE;

  // issue the loads first; everything below until the shared-memory store is
  // independent of them
  const int4* src = reinterpret_cast<const int4*>(delta + tbase) + w * 128 + lane;
  const int4 v0 = ld_stream_v4(src);
  const int4 v1 = ld_stream_v4(src + 32);
  const int4 v2 = ld_stream_v4(src + 64);
  const int4 v3 = ld_stream_v4(src + 96);

  const int c = L.blk2chrom[tbase >> GR_BLOCK_SHIFT];
  const u64 off = L.off[c];
  const u32 len = L.len[c];
  const bool act = (L.flags[c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
  const u32 jb = (u32)(tbase - off);                 // chromosome position of the tile's first cell
  const bool interior = jb >= 1 && (u64)jb + GR_SCAN_TILE <= (u64)len;

  int4* sw = sm_x + w * SCAN_PAD_INT4;
  { int g = lane;      sw[g + (g >> 2)] = v0; }
  { int g = lane + 32; sw[g + (g >> 2)] = v1; }
  { int g = lane + 64; sw[g + (g >> 2)] = v2; }
  { int g = lane + 96; sw[g + (g >> 2)] = v3; }
  __syncwarp();
  int d[SCAN_ITEMS];
  {
    const int4 x0 = sw[5 * lane + 0], x1 = sw[5 * lane + 1], x2 = sw[5 * lane + 2], x3 = sw[5 * lane + 3];
    d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w;
    d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
    d[8] = x2.x; d[9] = x2.y; d[10] = x2.z; d[11] = x2.w;
    d[12] = x3.x; d[13] = x3.y; d[14] = x3.z; d[15] = x3.w;
  }

  // thread-local inclusive sums and break mask
  u32 s[SCAN_ITEMS];
  u32 run = 0;
  u32 m = 0;
  const u32 j0 = jb + w * 512 + lane * 16;
  if (interior) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
      run += (u32)d[i];
      s[i] = run;
      m |= (d[i] != 0 ? 1u : 0u) << i;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
      run += (u32)d[i];
      s[i] = run;
      const u32 j = j0 + i;
      const bool b = (j == len) || (d[i] != 0 && j >= 1 && j < len);
      m |= (b ? 1u : 0u) << i;
    }
  }
  if (!act) m = 0;
  const u32 cnt = __popc(m);

  // warp scan of (sum, count); block combine
  const u32 wi_sum = warp_incl_scan_u32(run, lane);
  const u32 wi_cnt = warp_incl_scan_u32(cnt, lane);
  if (lane == 31) { sm_wsum[w] = wi_sum; sm_wcnt[w] = wi_cnt; }
  __syncthreads();
  u32 wx_sum = 0, wx_cnt = 0, t_sum = 0, t_cnt = 0;
#pragma unroll
  for (int k = 0; k < SCAN_WARPS; k++) {
    const u32 a = sm_wsum[k], b = sm_wcnt[k];
    if (k < w) { wx_sum += a; wx_cnt += b; }
    t_sum += a; t_cnt += b;
  }

  if (w == 0) {
    i64 agg[2] = { (i64)t_sum, (i64)t_cnt }, ex[2];
    lookback_exclusive<2>(lb, tile, agg, ex);
    if (lane == 0) {
      // keep the published sum inside 32 bits so the 62-bit payload never wraps
      sm_excl_sum = (u32)(u64)ex[0];
      sm_excl_cnt = (u64)ex[1];
    }
  }
  __syncthreads();
  const u32 ex_sum = sm_excl_sum;
  const u64 ex_cnt = sm_excl_cnt;

  if (tid == 0) {
    if (tbase == off) {
      out.chrom_start[c] = ex_cnt;
      if (ex_sum != 0) atomicOr(err, GR_DE_TAIL);   // previous chromosome did not return to 0 (2283-2289)
    }
    if (tile == ntiles - 1) {
      *out.total = ex_cnt + t_cnt;
      out.chrom_start[L.nchrom] = ex_cnt + t_cnt;
    }
  }

  // bitmap: 16 flags per thread, two lanes per 32-bit word
  {
    const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
    if (!(lane & 1)) bitmap[(tbase >> 5) + w * 16 + (lane >> 1)] = m | (hi << 16);
  }

  // emit the breaks of this thread
  if (m) {
    const u32 base = ex_sum + wx_sum + (wi_sum - run);          // exclusive prefix before d[0]
    u64 rank = ex_cnt + wx_cnt + (wi_cnt - cnt);
    bool neg = false;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
      if (m & (1u << i)) {
        const int N = (int)(base + (i ? s[i - 1] : 0u));
        neg |= N < 0;
        out.end[rank] = j0 + i;
        out.val[rank] = units_to_val(N < 0 ? 0 : N);
        rank++;
      }
    }
    if (neg) atomicOr(err, GR_DE_PILE);                          // ERRPILE 1921, 1969
  }
}

__global__ void k_fill_chrom_start(DevLayout L, u64* chrom_start, const u64* total) {
  if (threadIdx.x || blockIdx.x) return;
  u64 next = *total;
  chrom_start[L.nchrom] = next;
  for (int c = L.nchrom - 1; c >= 0; c--) {
    if (L.off[c] == ~0ull) chrom_start[c] = next;
    else next = chrom_start[c];
  }
}

void launch_fill_chrom_start(cudaStream_t s, const DevLayout& L, u64* chrom_start, const u64* total) {
  k_fill_chrom_start<<<1, 32, 0, s>>>(L, chrom_start, total); GR_NOTE_LAUNCH();
}

void launch_dense_scan(cudaStream_t s, const DevLayout& L, const int32_t* delta,
                       const ScanScratch& sc, DevRle out, u32* bitmap, int* err) {
  const u64 ntiles = L.T / GR_SCAN_TILE;
  cudaMemsetAsync(sc.st_sum, 0, ntiles * sizeof(u64), s);
  cudaMemsetAsync(sc.st_cnt, 0, ntiles * sizeof(u64), s);
  cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s);
  Lookback<2> lb;
  lb.st[0] = sc.st_sum; lb.st[1] = sc.st_cnt; lb.ticket = sc.ticket;
  k_dense_scan<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(delta, L, lb, out, bitmap, err, (u32)ntiles); GR_NOTE_LAUNCH();
  launch_fill_chrom_start(s, L, out.chrom_start, out.total);
}

// ============================================================================
// K2b: per chromosome, sum over its RLE intervals of (float)(end-start)*val
// (the float product of Genrich.c:2246 / 2018, accumulated there in a double).
// Here every float product is added EXACTLY in fixed point (integer part and
// 2^-40 fraction in separate u64 counters), so the result does not depend on the
// order of the atomics; the host rounds int + frac*2^-40 to a double once.
__global__ void __launch_bounds__(256)
k_rle_moment(DevRle r, int nchrom, u64* __restrict__ acc_int, u64* __restrict__ acc_frac) {
  __shared__ int sm_c0, sm_c1;
  __shared__ u64 sm_i[8], sm_f[8];
  const u64 n = *r.total;
  const u64 base = (u64)blockIdx.x * blockDim.x;
  if (base >= n) return;
  const u64 last = min(base + blockDim.x, n) - 1;
  if (threadIdx.x == 0) {
    sm_c0 = chrom_of_index(r.chrom_start, nchrom, base);
    sm_c1 = chrom_of_index(r.chrom_start, nchrom, last);
  }
  __syncthreads();
  const int c0 = sm_c0, c1 = sm_c1;
  const u64 i = base + threadIdx.x;
  u64 pi = 0, pf = 0;
  int c = c0;
  if (i < n) {
    if (c0 != c1) c = chrom_of_index(r.chrom_start, nchrom, i);
    const u32 e = r.end[i];
    const u32 st = (i == r.chrom_start[c]) ? 0u : r.end[i - 1];
    const float p = __fmul_rn(__uint2float_rn(e - st), r.val[i]);
    pi = (u64)p;                                     // p >= 0
    const float fr = __fsub_rn(p, (float)pi);        // exact: < 1 only when p < 2^24
    pf = (u64)(fr * 1099511627776.0f);               // * 2^40, exact
  }
  if (c0 == c1) {
    pi = warp_sum_u64(pi);
    pf = warp_sum_u64(pf);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sm_i[w] = pi; sm_f[w] = pf; }
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 ti = 0, tf = 0;
      for (int k = 0; k < 8; k++) { ti += sm_i[k]; tf += sm_f[k]; }
      ti += tf >> 40;
      tf &= (1ull << 40) - 1;
      if (ti) atomicAdd(acc_int + c0, ti);
      if (tf) atomicAdd(acc_frac + c0, tf);
    }
  } else if (i < n) {
    if (pi) atomicAdd(acc_int + c, pi);
    if (pf) atomicAdd(acc_frac + c, pf);
  }
}

void launch_rle_moment(cudaStream_t s, const DevRle& r, u64 n_upper, int nchrom,
                       u64* acc_int, u64* acc_frac) {
  if (!n_upper) return;
  const u64 blocks = (n_upper + 255) / 256;
  k_rle_moment<<<(unsigned)blocks, 256, 0, s>>>(r, nchrom, acc_int, acc_frac); GR_NOTE_LAUNCH();
}

__global__ void k_fill_u64(u64* p, u64 v, u64 n) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_u64(cudaStream_t s, u64* p, u64 v, u64 n) {
  if (n) { k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, v, n); GR_NOTE_LAUNCH(); }
}

u64 lookback_tiles_for(u64 n_items, u32 tile) { return (n_items + tile - 1) / tile; }
