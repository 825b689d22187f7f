// gr_common.cuh -- device-side building blocks shared by every kernel of
// libgenrich_cuda (sm_100a only).
//
//  * decoupled look-back tile prefix over K 62-bit counters (single-pass scan /
//    stream compaction: every byte of the input is read exactly once)
//  * warp / block scans, streaming 128-bit loads, relaxed gpu-scope status words
//  * the 1/120-unit -> reference float reconstruction (getVal, Genrich.c:1902)
#pragma once
#ifndef GR_EMU                      // tests/emu: the kernels compiled for the host (lock-step emulation)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

#define GR_FULL 0xffffffffu

// Slot layout: every owned chromosome occupies len+1 consecutive int32 delta
// cells starting at a multiple of GR_BLOCK_SLOTS, so that a dense tile, a bitmap
// block and a look-back tile never straddle two chromosomes.
#define GR_BLOCK_SLOTS 8192         // slots per bitmap block (256 x u32 words)
#define GR_BLOCK_SHIFT 13
#define GR_SCAN_TILE   4096         // slots per dense scan tile (256 thr x 16)

// chromosome flag bits (device copy of Chrom.skip/save + ownership)
#define GR_CF_OWNED 1               // owned by this context and not skipped
#define GR_CF_SAVE  2               // Chrom.save for the current replicate

// device error bits (OR-ed into one int)
#define GR_DE_POS    1
#define GR_DE_COUNT  2
#define GR_DE_CHROM  4
#define GR_DE_PILE   8
#define GR_DE_TAIL   16             // running sum not 0 at a chromosome start
#define GR_DE_TABLE  32             // hash table over its load limit (host retries)
#define GR_DE_CAP    64             // an optimistically sized buffer was too small (host retries)
#define GR_DE_EXPT   128            // no analyzable fragments in the experimental sample (2292)
#define GR_DE_SAT    256            // a delta cell beyond the reference's int16 range (saveInterval 2558-2573)

// A delta cell, in 1/120 units, that the reference's (int16 cov, uint8 frac) cell cannot hold.  The reference never
// gets there: it drops intervals once cov sits at INT16_MAX / INT16_MIN (saveInterval 2558-2573), and so does the
// fused path (k_sat_resolve, gr_dense.cu) -- there the test is an assertion.  The dense formulation (GR_FUSED=0)
// does not replay that arrival-order rule: it reports a cell that needed it (net deltas only: a conservative test).
#define GR_SAT_HI (32767 * 120 + 193)      /* cov 32767 plus the largest fraction (7/8 + 2/6 + 4/10) */
#define GR_SAT_LO (-32768 * 120)
__device__ __forceinline__ bool cell_saturated(int d) { return d > GR_SAT_HI || d < GR_SAT_LO; }
__device__ __forceinline__ bool cell_saturated_dense(int d) { return d >= 32768 * 120 || d < GR_SAT_LO; }

// ---------------------------------------------------------------------------
// memory helpers
#ifdef GR_EMU
__device__ __forceinline__ int4 ld_stream_v4(const int4* p) { return *p; }
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) { return *(const volatile u64*)p; }
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) { *(volatile u64*)p = v; }
#else
__device__ __forceinline__ int4 ld_stream_v4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
#endif

// ---------------------------------------------------------------------------
// warp primitives
__device__ __forceinline__ int warp_incl_scan_i32(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(GR_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ u32 warp_incl_scan_u32(u32 v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(GR_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ i64 warp_sum_i64(i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GR_FULL, v, o);
  return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GR_FULL, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// Decoupled look-back over K counters.  Each tile owns K status words; a word
// is flag(2 bits) | payload(62 bits, two's complement).  Flag 1: the payload is
// the tile's own aggregate; flag 2: the inclusive prefix up to and including the
// tile.  Every word is self-contained, so relaxed 64-bit accesses suffice.
// Tiles take their index from an atomic ticket, which guarantees that every
// predecessor of a running tile has started (forward progress).
#define GR_LB_PAYLOAD ((1ull << 62) - 1)
__device__ __forceinline__ u64 lb_pack(u64 flag, i64 v) { return (flag << 62) | ((u64)v & GR_LB_PAYLOAD); }
__device__ __forceinline__ i64 lb_val(u64 w) { return ((i64)(w << 2)) >> 2; }

template <int K>
struct Lookback {
  u64* st[K];      // each: ntiles words, zeroed before the launch
  u32* ticket;     // zeroed before the launch
};

// Called by all 32 lanes of one warp.  agg[k]: this tile's aggregate (identical
// in every lane).  Returns the exclusive prefix in excl[k] (all lanes).
// Each round inspects 32*PER predecessors (PER per lane, nearest first): with
// ~10^5 tiles retiring per millisecond a 32-wide window never reaches a tile
// whose inclusive prefix is already published, and the walk costs one L2 round
// trip per 32 tiles (measured on B200: 0.9 TB/s for the dense scan).
template <int K, int PER = 1>
__device__ __forceinline__ void lookback_exclusive(const Lookback<K>& lb, u32 tile,
                                                   const i64 (&agg)[K], i64 (&excl)[K]) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; k++) st_relaxed_u64(lb.st[k], lb_pack(2, agg[k]));
    }
#pragma unroll
    for (int k = 0; k < K; k++) excl[k] = 0;
    return;
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) st_relaxed_u64(lb.st[k] + tile, lb_pack(1, agg[k]));
  }
  i64 run[K];
#pragma unroll
  for (int k = 0; k < K; k++) run[k] = 0;
  i64 base = (i64)tile - 1;
  for (;;) {
    i64 part[K];
    u32 fmask;
    for (;;) {
      // all status loads of the round are issued before any is looked at
      u64 wv[PER][K];
#pragma unroll
      for (int e = 0; e < PER; e++) {
        const i64 idx = base - (i64)lane * PER - e;
#pragma unroll
        for (int k = 0; k < K; k++) {
          wv[e][k] = ld_relaxed_u64(lb.st[k] + (idx >= 0 ? idx : 0));
          if (idx < 0) wv[e][k] = lb_pack(2, 0);     // virtual tile before tile 0: inclusive prefix 0
        }
      }
      bool found = false, ok = true;
#pragma unroll
      for (int k = 0; k < K; k++) part[k] = 0;
#pragma unroll
      for (int e = 0; e < PER; e++) {
        const u64 f0 = wv[e][0] >> 62;
        bool valid = f0 != 0;
#pragma unroll
        for (int k = 1; k < K; k++) valid = valid && ((wv[e][k] >> 62) == f0);
        const bool use = !found;
        ok = ok && (!use || valid);
        if (use && valid) {
#pragma unroll
          for (int k = 0; k < K; k++) part[k] += lb_val(wv[e][k]);
          found = f0 == 2;
        }
      }
      fmask = __ballot_sync(GR_FULL, found);
      const int first = fmask ? (__ffs(fmask) - 1) : 32;
      const u32 need = first >= 31 ? GR_FULL : ((2u << first) - 1);
      const u32 bad = __ballot_sync(GR_FULL, !ok) & need;
      if (!bad) {
        if (lane > first) {
#pragma unroll
          for (int k = 0; k < K; k++) part[k] = 0;
        }
        break;
      }
    }
#pragma unroll
    for (int k = 0; k < K; k++) run[k] += warp_sum_i64(part[k]);
    if (fmask) break;
    base -= 32 * PER;
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) st_relaxed_u64(lb.st[k] + tile, lb_pack(2, run[k] + agg[k]));
  }
#pragma unroll
  for (int k = 0; k < K; k++) excl[k] = run[k];
}

// ---------------------------------------------------------------------------
// Pileup height in integer 1/120ths -> the reference's float.
// State of updateVal (Genrich.c:1915-1973) is (cov, frac = tenths<<5|sixths<<3|
// eighths) with value cov + e/8 + s/6 + t/10 (e in 0..7, s in 0..2, t in 0..4);
// N = 120cov + 15e + 20s + 12t has exactly one solution.  getVal (1902-1907)
// adds the four float terms left to right.
__device__ __forceinline__ float units_to_val(int N) {
  const int q = N / 120;
  if (N - q * 120 == 0) return (float)q;
  const int s = (2 * (N % 3)) % 3;
  const int t = (3 * (N % 5)) % 5;
  const int e = (4 * s + 4 * t - N) & 7;
  const int cov = (N - 15 * e - 20 * s - 12 * t) / 120;
  float v = __fadd_rn((float)cov, __fdiv_rn((float)e, 8.0f));
  v = __fadd_rn(v, __fdiv_rn((float)s, 6.0f));
  v = __fadd_rn(v, __fdiv_rn((float)t, 10.0f));
  return v;
}

// Table form for kernels that convert many heights: entry r = N mod 120 holds the three
// fractional terms (already divided, correctly rounded) and the borrow of the integer part.
__device__ __forceinline__ void units_lut_fill(float4* lut, int tid, int nthreads) {
  for (int r = tid; r < 120; r += nthreads) {
    const int s = (2 * (r % 3)) % 3;
    const int t = (3 * (r % 5)) % 5;
    const int e = (4 * s + 4 * t - r) & 7;
    const int borrow = (15 * e + 20 * s + 12 * t - r) / 120;       // 0 or 1
    lut[r] = make_float4(__fdiv_rn((float)e, 8.0f), __fdiv_rn((float)s, 6.0f), __fdiv_rn((float)t, 10.0f),
                         __int_as_float(borrow));
  }
}
__device__ __forceinline__ float units_to_val_lut(const float4* lut, int N) {
  const int q = N / 120;
  const float4 f = lut[N - q * 120];
  float v = __fadd_rn((float)(q - __float_as_int(f.w)), f.x);
  v = __fadd_rn(v, f.y);
  return __fadd_rn(v, f.z);
}

// index of the chromosome whose interval range [start[c], start[c+1]) holds i
// (start has n+1 monotone entries)
__device__ __forceinline__ int chrom_of_index(const u64* __restrict__ start, int n, u64 i) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}
