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
// K2: one pass over the dense int32 delta cells -- persistent, software-pipelined.
//
//   grid = resident CTAs only (2 per SM x 512 threads); each CTA loops over
//   8192-cell tiles handed out by an atomic ticket.  A tile is brought in with
//   cp.async (LDGSTS, 16 B per thread per op, coalesced) into a padded
//   shared-memory stage while the previous tile is being scanned, so two tiles
//   (64 KB) per CTA are always in flight towards HBM.  Each thread then owns 16
//   consecutive cells: thread / warp / block scan of (sum, #breaks), decoupled
//   look-back for the exclusive prefix of both, breaks written as (end, value) at
//   their global rank, 1 bit per cell into the break bitmap.
//
//   Look-back: one 128-bit status per tile {flag|sum32, flag|count}, written and
//   read with single 128-bit accesses.  The look-back warp inspects 256
//   predecessors per round (8 per lane): with ~10^5 tiles/ms retiring, a 32-wide
//   window never reaches a tile whose inclusive prefix is already published and
//   the chain degenerates to one L2 round trip per 32 tiles (measured: 0.9 TB/s).
//
// A break closes an interval at chromosome position j iff 1 <= j < len and
// delta[j] != 0, or j == len (Genrich.c:2241, 2268); its value is the running sum
// BEFORE delta[j] is added (2245), rebuilt as the reference float.
// Running sums are kept modulo 2^32: every true prefix fits in int32.
#define SCAN_THREADS 512
#define SCAN_WARPS 16
#define SCAN_ITEMS 16
#define SCAN_STAGE_INT4 2560      // 2048 int4 per tile + 1 pad per 4
#define SCAN_STAGES 2
#define SCAN_LB_PER_LANE 10

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_status(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(ulonglong2* p, u64 flag, u32 sum, u64 cnt) {
  const u64 a = (flag << 62) | sum, b = (flag << 62) | cnt;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// exclusive (sum, count) prefix of `tile`; called by the 32 lanes of one warp
__device__ __forceinline__ void scan_lookback(ulonglong2* __restrict__ st, u32 tile, u32 agg_sum,
                                              u32 agg_cnt, u32& ex_sum, u64& ex_cnt) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_status(st, 2, agg_sum, agg_cnt);
    ex_sum = 0; ex_cnt = 0;
    return;
  }
  if (lane == 0) st_status(st + tile, 1, agg_sum, agg_cnt);
  u32 run_s = 0;
  u64 run_c = 0;
  i64 base = (i64)tile - 1;
  for (;;) {
    u32 ls = 0; u64 lc = 0;
    bool found, ok;
    do {
      ls = 0; lc = 0; found = false; ok = true;
#pragma unroll
      for (int e = 0; e < SCAN_LB_PER_LANE; e++) {
        const i64 idx = base - (i64)lane * SCAN_LB_PER_LANE - e;
        ulonglong2 w;
        if (idx >= 0) w = ld_status(st + idx);
        else { w.x = 2ull << 62; w.y = 2ull << 62; }        // before tile 0: inclusive prefix 0
        const u64 f = w.x >> 62;
        const bool valid = f != 0 && (w.y >> 62) == f;
        if (!found) {
          if (!valid) ok = false;
          else {
            ls += (u32)w.x;
            lc += w.y & GR_LB_PAYLOAD;
            if (f == 2) found = true;
          }
        }
      }
      // lanes beyond the first lane holding an inclusive prefix do not matter
      const u32 fmask = __ballot_sync(GR_FULL, found);
      const int first = fmask ? (__ffs(fmask) - 1) : 32;
      const u32 bad = __ballot_sync(GR_FULL, !ok) & (first >= 31 ? GR_FULL : ((2u << first) - 1));
      if (!bad) {
        if (lane > first) { ls = 0; lc = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ls += __shfl_xor_sync(GR_FULL, ls, o);
          lc += __shfl_xor_sync(GR_FULL, lc, o);
        }
        run_s += ls; run_c += lc;
        found = fmask != 0;
        break;
      }
    } while (true);
    if (found) break;
    base -= 32 * SCAN_LB_PER_LANE;
  }
  if (lane == 0) st_status(st + tile, 2, run_s + agg_sum, run_c + agg_cnt);
  ex_sum = run_s; ex_cnt = run_c;
}

struct TileMeta { u64 off; u32 len; int c; bool act; };
__device__ __forceinline__ TileMeta tile_meta(const DevLayout& L, u32 tile, u32 ntiles) {
  TileMeta m;
  m.c = 0; m.off = 0; m.len = 0; m.act = false;
  if (tile < ntiles) {
    m.c = L.blk2chrom[tile];
    m.off = L.off[m.c];
    m.len = L.len[m.c];
    m.act = (L.flags[m.c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
  }
  return m;
}

// Tiles are dealt round-robin: CTA b takes tiles b, b+G, b+2G, ... (G = gridDim.x,
// all CTAs co-resident: cooperative launch).  In round k every CTA works on a tile
// of [kG, (k+1)G), so a tile only ever waits for aggregates of tiles of its own
// round and finds the inclusive prefixes of the previous round within one
// 320-wide look-back window; handing tiles out through an atomic ticket that is
// taken early enough to prefetch makes low tickets wait behind high ones
// (measured: 6x slower than no pipelining at all).
__global__ void __launch_bounds__(SCAN_THREADS, 2)
k_dense_scan(const int32_t* __restrict__ delta, DevLayout L, ulonglong2* __restrict__ status,
             DevRle out, u32* __restrict__ bitmap, int* __restrict__ err, u32 ntiles) {
  extern __shared__ int4 sm_x[];                       // SCAN_STAGES * SCAN_STAGE_INT4
  __shared__ u32 sm_wsum[SCAN_WARPS];
  __shared__ u32 sm_wcnt[SCAN_WARPS];
  __shared__ u32 sm_excl_sum;
  __shared__ u64 sm_excl_cnt;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u32 G = gridDim.x;

  // issue the loads of one tile into a stage: chunk g (16 B) -> slot g + g/4
  auto issue = [&](u32 tile, int stage) {
    if (tile < ntiles) {
      const int4* src = reinterpret_cast<const int4*>(delta + (u64)tile * GR_BLOCK_SLOTS);
      int4* dst = sm_x + stage * SCAN_STAGE_INT4;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int g = k * SCAN_THREADS + tid;
        cp_async16(dst + g + (g >> 2), src + g);
      }
    }
    cp_async_commit();
  };

  u32 tile = blockIdx.x;
  issue(tile, 0);
  issue(tile + G, 1);
  TileMeta meta = tile_meta(L, tile, ntiles);

  for (u32 it = 0; tile < ntiles; it++, tile += G) {
    const int stage = it & 1;
    cp_async_wait<1>();
    __syncthreads();                                   // this tile's cells are in shared memory

    int d[SCAN_ITEMS];
    {
      const int4* sx = sm_x + stage * SCAN_STAGE_INT4 + 5 * tid;
      const int4 x0 = sx[0], x1 = sx[1], x2 = sx[2], x3 = sx[3];
      d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w;
      d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
      d[8] = x2.x; d[9] = x2.y; d[10] = x2.z; d[11] = x2.w;
      d[12] = x3.x; d[13] = x3.y; d[14] = x3.z; d[15] = x3.w;
    }
    __syncthreads();                                   // stage is free again
    issue(tile + 2 * G, stage);                        // two rounds ahead -> this stage
    const TileMeta meta_next = tile_meta(L, tile + G, ntiles);   // consumed next round

    const u64 tbase = (u64)tile * GR_BLOCK_SLOTS;
    const int c = meta.c;
    const u64 off = meta.off;
    const u32 len = meta.len;
    const bool act = meta.act;
    const u32 jb = (u32)(tbase - off);                 // chromosome position of the tile's first cell
    const bool interior = jb >= 1 && (u64)jb + GR_BLOCK_SLOTS <= (u64)len;

    // thread-local inclusive sums and break mask
    u32 s[SCAN_ITEMS];
    u32 run = 0, m = 0;
    const u32 j0 = jb + tid * SCAN_ITEMS;
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
      u32 es; u64 ec;
      scan_lookback(status, tile, t_sum, t_cnt, es, ec);
      if (lane == 0) { sm_excl_sum = es; sm_excl_cnt = ec; }
    }
    __syncthreads();
    const u32 ex_sum = sm_excl_sum;
    const u64 ex_cnt = sm_excl_cnt;

    if (tid == 0) {
      if (tbase == off) {
        out.chrom_start[c] = ex_cnt;
        if (ex_sum != 0) atomicOr(err, GR_DE_TAIL);    // previous chromosome did not return to 0 (2283-2289)
      }
      if (tile == ntiles - 1) {
        *out.total = ex_cnt + t_cnt;
        out.chrom_start[L.nchrom] = ex_cnt + t_cnt;
      }
    }
    {
      const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
      if (!(lane & 1)) bitmap[(tbase >> 5) + (tid >> 1)] = m | (hi << 16);
    }
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
    meta = meta_next;
    // sm_wsum / sm_excl are rewritten only after the next round's first barrier
  }
  cp_async_wait<0>();
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
  const u64 ntiles = L.nblocks;                        // one tile per 8192-cell block
  static int grid = 0;
  const size_t smem = (size_t)SCAN_STAGES * SCAN_STAGE_INT4 * sizeof(int4);
  if (!grid) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_dense_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_dense_scan, SCAN_THREADS, smem);
    if (per < 1) per = 1;
    if (per > 2) per = 2;
    grid = sms * per;                                  // persistent: co-resident CTAs only
  }
  cudaMemsetAsync(sc.st_sum, 0, ntiles * sizeof(ulonglong2), s);
  unsigned g = (unsigned)(ntiles < (u64)grid ? ntiles : (u64)grid);
  ulonglong2* st = (ulonglong2*)sc.st_sum;
  u32 nt = (u32)ntiles;
  DevLayout Lc = L;
  void* args[] = { (void*)&delta, (void*)&Lc, (void*)&st, (void*)&out, (void*)&bitmap, (void*)&err, (void*)&nt };
  // cooperative launch: fails instead of deadlocking if the CTAs cannot all be resident
  cudaLaunchCooperativeKernel((const void*)k_dense_scan, dim3(g), dim3(SCAN_THREADS), args, smem, s);
  GR_NOTE_LAUNCH();
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
