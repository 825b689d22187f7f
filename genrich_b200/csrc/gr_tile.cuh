// gr_tile.cuh -- block-level helpers for the interval-domain kernels: ticketed
// tiles, compaction ranks through the look-back, and "which chromosome does this
// run of interval indices belong to".
#pragma once
#include "gr_common.cuh"

__device__ __forceinline__ u32 take_ticket(u32* ticket) {
  __shared__ u32 sm_ticket;
  if (threadIdx.x == 0) sm_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  return sm_ticket;
}

// Exclusive global rank of this thread's first flagged item, given its flagged
// count.  All threads of the block call it once.  blockDim.x <= 1024.
__device__ __forceinline__ u64 tile_exclusive_rank(const Lookback<1>& lb, u32 tile, u32 cnt,
                                                   u32& tile_total) {
  __shared__ u32 sm_w[32];
  __shared__ u64 sm_ex;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const u32 wi = warp_incl_scan_u32(cnt, lane);
  if (lane == 31) sm_w[w] = wi;
  __syncthreads();
  u32 wx = 0, tot = 0;
  for (int k = 0; k < nw; k++) {
    const u32 a = sm_w[k];
    if (k < w) wx += a;
    tot += a;
  }
  if (w == 0) {
    i64 agg[1] = { (i64)tot }, ex[1];
    lookback_exclusive<1>(lb, tile, agg, ex);
    if (lane == 0) sm_ex = (u64)ex[0];
  }
  __syncthreads();
  tile_total = tot;
  return sm_ex + wx + (wi - cnt);
}

// Chromosomes of the first and last interval index of a tile (one binary search
// each by thread 0).  When they agree no chromosome boundary lies inside.
struct TileChrom { int c0, c1; };
__device__ __forceinline__ TileChrom tile_chrom_range(const u64* __restrict__ chrom_start,
                                                      int nchrom, u64 first, u64 last) {
  __shared__ int sm_c[2];
  if (threadIdx.x == 0) {
    sm_c[0] = chrom_of_index(chrom_start, nchrom, first);
    sm_c[1] = chrom_of_index(chrom_start, nchrom, last);
  }
  __syncthreads();
  TileChrom t;
  t.c0 = sm_c[0];
  t.c1 = sm_c[1];
  return t;
}
