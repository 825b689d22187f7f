// gr_dense.cu -- the per-base kernels: K1 delta scatter, K2 single-pass dense
// prefix sum with break compaction, K2b exact weighted-length reduction.
//
// Replaces saveInterval's diff-array writes (Genrich.c:2575-2583) and the
// sequential O(genome) loops of savePileupExpt (2239-2273) / calcFactor
// (2013-2038).  HBM-bound int32 work: no tensor cores.
#include "gr_common.cuh"
#include "gr_internal.h"
#include <stdlib.h>
#include <math.h>
#include <stdio.h>

// Function attributes (the dynamic shared memory limit) are per DEVICE: a process that drives several contexts
// (genrich-b200 --gpus N) sets them on each device it launches the kernel on.  which: one small number per kernel.
static bool first_use_on_device(int which) {
  static unsigned long long seen[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (seen[which & 7] >> dev & 1ull) return false;
  seen[which & 7] |= 1ull << dev;
  return true;
}

// ============================================================================
// K1: two int32 reductions (RED.ADD) per interval record into the dense delta
// array, in units of 1/120 (weights 120/count, count in {1,2,3,4,5,6,8,10}:
// addFrac 2311 / subFrac 2412).  Clamping as saveInterval 2522-2544.
// Two forms.  Small samples: k_scatter, the two reductions straight into the (all-zero)
// array.  Large samples: the records are first bucketed by the 8192-cell block of their
// start (count -> scan -> cursor move; the bucket array is written once, 8 B per record),
// then k_sb_build assembles every block of the delta array in shared memory and writes it
// out densely -- the array is WRITTEN sequentially once instead of being hit by 2 random
// read-modify-writes per record (ncu: 8.5 GB read + 3.2 GB written per 50 M records, 30 %
// of DRAM peak), and needs no clearing before or after.

// One record, decoded and checked as saveInterval does (2522-2544).  PACKED: 8-byte records
// (include/genrich_cuda.h, GR_PACK), else int32 x 4.  Returns false for records that are
// dropped (unsaved chromosome) or in error (flagged in e_local).
// raw record: the 8-byte packed word (in .x/.y) or the int32 x 4 form
template <bool PACKED>
__device__ __forceinline__ int4 load_raw(const void* __restrict__ recs, u64 i) {
  if (PACKED) {
    const u64 v = __ldcs(reinterpret_cast<const u64*>(recs) + i);
    return make_int4((int)(u32)v, (int)(u32)(v >> 32), 0, 0);
  }
  return ld_stream_v4(reinterpret_cast<const int4*>(recs) + i);
}
template <bool PACKED>
__device__ __forceinline__ bool decode_raw(const int4 r, const DevLayout& L,
                                           u64& s_slot, u32& span, int& w, int& e_local, u32& c_local) {
  int c, cnt;
  i64 s, e;
  if (PACKED) {
    const u32 hi = (u32)r.y;
    s = (i64)(u32)r.x;
    e = s + (i64)(hi & 0x3fffu);
    c = (int)((hi >> 14) & 0x3fffu);
    cnt = (int)(hi >> 28);
  } else {
    c = r.x; s = r.y; e = r.z; cnt = r.w;
  }
  if (c < 0 || c >= L.nchrom) { e_local |= GR_DE_CHROM; return false; }
  const uint8_t f = L.flags[c];
  const u64 off = L.off[c];
  if (off == ~0ull) {                      // not owned by this context, or -e skipped
    if (!(f & GR_CF_OWNED)) e_local |= GR_DE_CHROM;
    return false;
  }
  if (!(f & GR_CF_SAVE)) return false;     // processPair 3137-3138: not in this replicate
  if (cnt < 1 || cnt > 10 || !((1 << cnt) & 0x57E)) { e_local |= GR_DE_COUNT; return false; }
  const i64 len = L.len[c];
  bool cl = false;
  if (s < 0) { s = 0; cl = true; }
  if (s >= len || e < s) { e_local |= GR_DE_POS; return false; }
  if (e > len) { e = len; cl = true; }
  c_local += cl;
  s_slot = off + (u64)s;
  span = (u32)(e - s);
  w = 120 / cnt;
  return true;
}
template <bool PACKED>
__device__ __forceinline__ bool decode_record(const void* __restrict__ recs, u64 i, const DevLayout& L,
                                              u64& s_slot, u32& span, int& w, int& e_local, u32& c_local) {
  return decode_raw<PACKED>(load_raw<PACKED>(recs, i), L, s_slot, span, w, e_local, c_local);
}

template <bool PACKED>
__global__ void __launch_bounds__(256)
k_scatter(const void* __restrict__ recs, u64 n, DevLayout L, int32_t* __restrict__ delta,
          int* __restrict__ err, u64* __restrict__ clamped) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u64 s_slot; u32 span; int w;
    if (!decode_record<PACKED>(recs, i, L, s_slot, span, w, e_local, c_local)) continue;
    atomicAdd(delta + s_slot, w);
    atomicAdd(delta + s_slot + span, -w);
  }
  if (e_local) atomicOr(err, e_local);
  if (c_local) atomicAdd(clamped, (u64)c_local);
}

void launch_scatter(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                    int32_t* delta, int* err, u64* clamped) {
  if (!n) return;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (packed) k_scatter<true><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, delta, err, clamped);
  else k_scatter<false><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, delta, err, clamped);
  GR_NOTE_LAUNCH();
}

// ---- bucketed build ------------------------------------------------------------------
// bucket entry (4 bytes): bits 0-14 interval length in cells, 15-27 offset of the start inside
// its block, 28-31 count.  Intervals of SB_LONG cells or more (none in sequencing data, but
// legal) bypass the buckets: both their ends go to the spill list, applied by reductions
// after the blocks are written -- as do the ends of intervals that reach into a later block.
// 6-byte records (GR_PACK6: cell of the start in this context's layout, length, count) -> the
// 8-byte GR_PACK words every other kernel reads.  One streaming pass: 6 B in, 8 B out per record.
__global__ void __launch_bounds__(256)
k_unpack6(const unsigned short* __restrict__ recs, u64 n, DevLayout L, u64* __restrict__ out, int* __restrict__ err) {
  const u64 stride = (u64)gridDim.x * 256;
  int e_local = 0;
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const u32 w0 = __ldcs(recs + 3 * i), w1 = __ldcs(recs + 3 * i + 1), w2 = __ldcs(recs + 3 * i + 2);
    const u64 cell = (u64)(w0 | (w1 << 16));
    u64 v;
    if (cell >= L.T) { e_local |= GR_DE_CHROM; v = 0; }      // count 0: dropped (and reported) by the next pass
    else {
      const int c = L.blk2chrom[cell >> GR_BLOCK_SHIFT];
      const u64 start = cell - L.off[c];                     // >= len (padding cells): ERRPOS in the next pass
      v = start | ((u64)(w2 & 0xfffu) << 32) | ((u64)c << 46) | ((u64)(w2 >> 12) << 60);
    }
    out[i] = v;
  }
  if (e_local) atomicOr(err, e_local);
}
void launch_unpack6(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, u64* out, int* err) {
  if (!n) return;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_unpack6<<<(unsigned)blocks, 256, 0, s>>>((const unsigned short*)recs, n, L, out, err); GR_NOTE_LAUNCH();
}

#define SB_LONG (1u << 15)
__device__ __forceinline__ uint2 sb_spill_entry(u64 slot, int w) {       // w signed, |w| <= 120
  return make_uint2((u32)slot, ((u32)(slot >> 32) << 8) | ((u32)w & 0xffu));
}
template <bool PACKED>
__global__ void __launch_bounds__(256)
k_sb_count(const void* __restrict__ recs, u64 n, DevLayout L, u32* __restrict__ blk_cnt,
           int* __restrict__ err, u64* __restrict__ clamped) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u64 s_slot; u32 span; int w;
    if (!decode_record<PACKED>(recs, i, L, s_slot, span, w, e_local, c_local)) continue;
    if (span < SB_LONG) atomicAdd(blk_cnt + (s_slot >> GR_BLOCK_SHIFT), 1u);
  }
  if (e_local) atomicOr(err, e_local);       // errors and clamp counts are reported by this pass only
  if (c_local) atomicAdd(clamped, (u64)c_local);
}

// exclusive scan of the per-block counts (~4e5 for a human genome) in three small steps:
// sums of 4096-counter chunks, scan of the <= 1024 chunk sums (one CTA), rescan with the base
#define SB_CHUNK 4096
// sat_flag (may be NULL): word 0 is raised when a block holds SAT_MIN_EVENTS entries or more -- only such a block
// can hold a cell the reference's int16 counters saturate on (k_sat_resolve below); word 1 collects the entries
// that lie in blocks of FORM_HOT_MIN entries or more (the scan form is chosen by it on the device: form_skip)
#define SAT_MIN_EVENTS 32767u
#define FORM_HOT_MIN 1024u
__global__ void __launch_bounds__(256)
k_sb_scan1(const u32* __restrict__ blk_cnt, u32* __restrict__ chunk_sum, u64 nblocks, u32* __restrict__ sat_flag) {
  __shared__ u32 sh[8];
  const u64 base = (u64)blockIdx.x * SB_CHUNK;
  __shared__ u32 sh_hot[8];
  u32 s = 0, hot = 0;
  bool big = false;
  for (int i = threadIdx.x; i < SB_CHUNK; i += 256)
    if (base + i < nblocks) {
      const u32 v = blk_cnt[base + i];
      s += v; big |= v >= SAT_MIN_EVENTS;
      if (v >= FORM_HOT_MIN) hot += v;
    }
  if (big && sat_flag) atomicOr(sat_flag, 1u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(GR_FULL, s, o); hot += __shfl_xor_sync(GR_FULL, hot, o); }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; sh_hot[threadIdx.x >> 5] = hot; }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0, h = 0;
    for (int k = 0; k < 8; k++) { t += sh[k]; h += sh_hot[k]; }
    chunk_sum[blockIdx.x] = t;
    if (h && sat_flag) atomicAdd(sat_flag + 1, h);
  }
}
__global__ void __launch_bounds__(1024)
k_sb_scan2(u32* __restrict__ chunk_sum, u32 nchunks, u32* __restrict__ blk_start, u64 nblocks) {
  __shared__ u32 sh[1024];
  const int t = threadIdx.x;
  u32 carry = 0;
  for (u32 c0 = 0; c0 < nchunks; c0 += 1024) {          // one round unless the genome has > 4e6 blocks
    const u32 v = c0 + t < nchunks ? chunk_sum[c0 + t] : 0u;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const u32 x = t >= o ? sh[t - o] : 0u;
      __syncthreads();
      sh[t] += x;
      __syncthreads();
    }
    if (c0 + t < nchunks) chunk_sum[c0 + t] = carry + sh[t] - v;    // exclusive
    const u32 tot = sh[1023];
    __syncthreads();
    carry += tot;
  }
  if (t == 0) blk_start[nblocks] = carry;
}
__global__ void __launch_bounds__(256)
k_sb_scan3(const u32* __restrict__ blk_cnt, const u32* __restrict__ chunk_base, u32* __restrict__ blk_start,
           u32* __restrict__ cursor, u64 nblocks) {
  __shared__ u32 sh[8];
  const u64 base = (u64)blockIdx.x * SB_CHUNK + threadIdx.x * 16;     // 16 consecutive counters per thread
  u32 v[16], s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { v[i] = base + i < nblocks ? blk_cnt[base + i] : 0u; s += v[i]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u32 wi = warp_incl_scan_u32(s, lane);
  if (lane == 31) sh[w] = wi;
  __syncthreads();
  u32 run = chunk_base[blockIdx.x] + wi - s;
  for (int k = 0; k < w; k++) run += sh[k];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (base + i < nblocks) { blk_start[base + i] = run; cursor[base + i] = run; }
    run += v[i];
  }
}

template <bool PACKED>
__global__ void __launch_bounds__(256)
k_sb_move(const void* __restrict__ recs, u64 n, DevLayout L, u32* __restrict__ cursor, u32* __restrict__ bucketed,
          uint2* __restrict__ spill, u32* __restrict__ spill_ctr) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u64 s_slot; u32 span; int w;
    if (!decode_record<PACKED>(recs, i, L, s_slot, span, w, e_local, c_local)) continue;
    if (span < SB_LONG) {
      const u32 pos = atomicAdd(cursor + (s_slot >> GR_BLOCK_SHIFT), 1u);
      bucketed[pos] = span | ((u32)(s_slot & (GR_BLOCK_SLOTS - 1)) << 15) | ((u32)(120 / w) << 28);
    } else {
      const u32 k = atomicAdd(spill_ctr, 2u);
      spill[k] = sb_spill_entry(s_slot, w);
      spill[k + 1] = sb_spill_entry(s_slot + span, -w);
    }
  }
}

// One CTA per 8192-cell block: assemble it in shared memory, write it out.  An interval
// that ends in a later block leaves its end to k_sb_spill (the later block is written by
// another CTA, at an unknown time).
__global__ void __launch_bounds__(256)
k_sb_build(const u32* __restrict__ bucketed, const u32* __restrict__ blk_start, int32_t* __restrict__ delta,
           uint2* __restrict__ spill, u32* __restrict__ spill_ctr) {
  __shared__ int4 sm4[GR_BLOCK_SLOTS / 4];
  int* sm = reinterpret_cast<int*>(sm4);
  const u32 blk = blockIdx.x;
  const u32 a = blk_start[blk], b = blk_start[blk + 1];
#pragma unroll
  for (int i = 0; i < 8; i++) sm4[i * 256 + threadIdx.x] = make_int4(0, 0, 0, 0);
  __syncthreads();
  for (u32 i = a + threadIdx.x; i < b; i += 256) {
    const u32 v = __ldcs(bucketed + i);
    const u32 so = (v >> 15) & (GR_BLOCK_SLOTS - 1), span = v & (SB_LONG - 1);
    const int w = 120 / (int)(v >> 28);
    atomicAdd(sm + so, w);
    const u32 eo = so + span;
    if (eo < GR_BLOCK_SLOTS) atomicAdd(sm + eo, -w);
    else spill[atomicAdd(spill_ctr, 1u)] = sb_spill_entry(((u64)blk << GR_BLOCK_SHIFT) + eo, -w);
  }
  __syncthreads();
  int4* out = reinterpret_cast<int4*>(delta) + (u64)blk * (GR_BLOCK_SLOTS / 4);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i * 256 + threadIdx.x] = sm4[i * 256 + threadIdx.x];
}

__global__ void __launch_bounds__(256)
k_sb_spill(const uint2* __restrict__ spill, const u32* __restrict__ spill_ctr, int32_t* __restrict__ delta) {
  const u32 n = *spill_ctr;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint2 v = spill[i];
    atomicAdd(delta + (((u64)(v.y >> 8) << 32) | v.x), (int)(signed char)(v.y & 0xffu));
  }
}

void launch_sb_count(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                     u32* blk_cnt, int* err, u64* clamped) {
  if (!n) return;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (packed) k_sb_count<true><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, blk_cnt, err, clamped);
  else k_sb_count<false><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, blk_cnt, err, clamped);
  GR_NOTE_LAUNCH();
}
// chunk_sum: scratch of ceil(nblocks / 4096) words
void launch_sb_scan(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* blk_start, u32* cursor, u32* chunk_sum) {
  launch_sb_scan_a(s, nbuckets, blk_cnt, chunk_sum, nullptr);
  launch_sb_scan_b(s, nbuckets, blk_cnt, blk_start, cursor, chunk_sum);
}
// the scan in two halves, so that k_sat_resolve can run in between (it may take entries out of blocks)
void launch_sb_scan_a(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* chunk_sum, u32* sat_flag) {
  const u32 nchunks = (u32)((nbuckets + SB_CHUNK - 1) / SB_CHUNK);
  k_sb_scan1<<<nchunks, 256, 0, s>>>(blk_cnt, chunk_sum, nbuckets, sat_flag); GR_NOTE_LAUNCH();
}
void launch_sb_scan_b(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* blk_start, u32* cursor, u32* chunk_sum) {
  const u32 nchunks = (u32)((nbuckets + SB_CHUNK - 1) / SB_CHUNK);
  k_sb_scan2<<<1, 1024, 0, s>>>(chunk_sum, nchunks, blk_start, nbuckets); GR_NOTE_LAUNCH();
  k_sb_scan3<<<nchunks, 256, 0, s>>>(blk_cnt, chunk_sum, blk_start, cursor, nbuckets); GR_NOTE_LAUNCH();
}
// spill_ctr must be zero before the first move of a sample
void launch_sb_move(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed, u32* cursor, u32* bucketed,
                    uint2* spill, u32* spill_ctr) {
  if (!n) return;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (packed) k_sb_move<true><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, cursor, bucketed, spill, spill_ctr);
  else k_sb_move<false><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, cursor, bucketed, spill, spill_ctr);
  GR_NOTE_LAUNCH();
}
void launch_sb_build(cudaStream_t s, const DevLayout& L, const u32* bucketed, const u32* blk_start,
                     int32_t* delta, uint2* spill, u32* spill_ctr) {
  k_sb_build<<<(unsigned)L.nblocks, 256, 0, s>>>(bucketed, blk_start, delta, spill, spill_ctr); GR_NOTE_LAUNCH();
  k_sb_spill<<<148 * 2, 256, 0, s>>>(spill, spill_ctr, delta); GR_NOTE_LAUNCH();
}

// ============================================================================
// K2: the per-base pass -- one read of the dense int32 delta cells (4 B per base per
// sample array: the algorithmic bytes of the roofline), prefix sum, break compaction.
//
// A break closes an interval at chromosome position j iff 1 <= j < len and
// delta[j] != 0, or j == len (Genrich.c:2241, 2268); its value is the running sum
// BEFORE delta[j] is added (2245), rebuilt as the reference float.
// Running sums are kept modulo 2^32: every true prefix fits in int32.
//
// History of the single-pass variants this replaced (profiles/README.md): decoupled
// look-back 0.92 TB/s; persistent round-robin tiles with a two-level aggregate exchange,
// an exchange warp and emission two tiles late 2.3 TB/s -- all of them limited by the
// prefix hand-over between CTAs, not by HBM.
#define SC_ITEMS 16

#ifdef GR_EMU                       // tests/emu: the copy completes at once (a stronger guarantee than the ring relies on)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
#endif

// blocked read of a thread's 16 cells from an (unswizzled) stage: used only by the
// rare dense path, so the 4-way bank conflict does not matter
__device__ __forceinline__ void sc_load_items(const int4* stage, int tid, int (&d)[SC_ITEMS]) {
  const int4 x0 = stage[4 * tid + 0], x1 = stage[4 * tid + 1];
  const int4 x2 = stage[4 * tid + 2], x3 = stage[4 * tid + 3];
  d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w;
  d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
  d[8] = x2.x; d[9] = x2.y; d[10] = x2.z; d[11] = x2.w;
  d[12] = x3.x; d[13] = x3.y; d[14] = x3.z; d[15] = x3.w;
}

// K2, streaming form: no cross-warp dependency at all inside the 4 B/cell pass.
//
// Why: in a single-pass chained scan a tile's breaks need the tile's global prefix
// (running height, rank) before they can be written, i.e. a store -> poll hop between CTAs.
// Measured on B200 under a 2 TB/s stream: one poll of the status words takes ~1.4 us and an
// exchange ~5 polls (everybody waits for the round's slowest CTA): 7-11 us against a tile
// time of 4 us; two or three tiles of slack did not hide it and 30-50 % of all warp samples
// sat on the prefix wait.
//
// So the pass over the cells does not wait for anything:
//   K2a k_scan_stream  every WARP owns a contiguous run of 512-cell spans and walks it
//        alone with its own cp.async ring, carrying (height, #breaks) in registers, both
//        relative to the start of its run.  Breaks are appended -- as (end coordinate,
//        run-relative height) -- to 256-entry pages taken from a global page counter; the
//        break bitmap is written and the non-zero cells are cleared on the way.
//   K2b k_scan_fix     one CTA: exclusive scan over the per-warp totals (<= 8192 warps),
//        chromosome starts, tail check.
//   K2c k_scan_place   moves every page to its final rank, adds the warp's base height and
//        rebuilds the reference float: 8 B read + 8 B written per INTERVAL (~0.1 B/cell).
#define SS_PAGE 256
#define SS_PAGE_SHIFT 8
#define SS_MAX_WARPS 8192
// pages beyond cap / SS_PAGE: two per warp of k_scan_stream (8192 x 2), or up to 66 per CTA of
// k_fb_scan (148 x 6 x 66 = 58608)
#define SS_SPARE_PAGES (16 * SS_MAX_WARPS)
struct StreamWs {
  uint2* pent;                 // provisional entries (end coordinate, run-relative height), max_pages * SS_PAGE
  uint2* page_meta;            // page -> (warp, sequence number inside the warp's run)
  u32* page_ctr;               // pages handed out
  uint2* warp_tot;             // per warp: (sum of its deltas, its #breaks)
  ulonglong2* warp_base;       // per warp: (height, rank) at the start of its run
  uint4* marks;                // per chromosome: (warp, height, rank) at its first cell, run-relative
  u32 max_pages;
};

template <int NSTAGE, int CPS>
__global__ void __launch_bounds__(512, CPS)
k_scan_stream(int32_t* __restrict__ delta, DevLayout L, StreamWs W, u32* __restrict__ bitmap,
              int* __restrict__ err, u32 nspans, u32 R, int zero_after) {
  extern __shared__ int4 sm_x[];                       // per warp NSTAGE stages of 128 chunks
  __shared__ u32 sm_bm[16 * 16];
  __shared__ unsigned char sm_list[16 * 128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u32 gw = blockIdx.x * 16 + w;
  const u32 s0 = gw * R, s1 = min(s0 + R, nspans);
  if (s0 >= s1) {
    if (lane == 0 && gw < SS_MAX_WARPS) W.warp_tot[gw] = make_uint2(0, 0);
    return;
  }
  int4* const wring = sm_x + w * (NSTAGE * 128) + lane;
  unsigned char* const wlist = sm_list + w * 128;
  u32* const wbm = sm_bm + w * 16;
  const int4* src = reinterpret_cast<const int4*>(delta) + (u64)s0 * 128 + lane;   // span being consumed
  auto issue = [&](bool on, int stage, const int4* from) {
    if (on) {
      int4* dst = wring + stage * 128;
#pragma unroll
      for (int r = 0; r < 4; r++) cp_async16(dst + r * 32, from + r * 32);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < NSTAGE; st++) issue(s0 + st < s1, st, src + (u64)st * 128);

  // output pages: `cur` is being filled (sequence cur_seq), `nxt` is in hand so that a batch
  // may run over the page end; the page after that is requested as soon as `nxt` becomes `cur`.
  // Page numbers are clamped to the last page: cannot happen (#breaks <= cap), never out of bounds.
  const u32 last_page = W.max_pages - 1;
  u32 cur = 0, nxt = 0, cur_seq = 0, pend_reg = 0;
  bool pending = false;
  if (lane == 0) {
    cur = atomicAdd(W.page_ctr, 2u);
    if (cur + 1 >= last_page) { atomicOr(err, GR_DE_TABLE); cur = last_page - 1; }
    W.page_meta[cur] = make_uint2(gw, 0); W.page_meta[cur + 1] = make_uint2(gw, 1);
  }
  cur = __shfl_sync(GR_FULL, cur, 0);
  nxt = cur + 1;
  auto resolve = [&]() {
    if (pending) {
      if (lane == 0) {
        if (pend_reg >= last_page) { atomicOr(err, GR_DE_TABLE); pend_reg = last_page; }
        W.page_meta[pend_reg] = make_uint2(gw, cur_seq + 1);
      }
      nxt = __shfl_sync(GR_FULL, pend_reg, 0);
      pending = false;
    }
  };
  auto advance = [&]() {                                // nxt becomes cur, ask for another page
    resolve();
    cur = nxt; cur_seq++;
    if (lane == 0) pend_reg = atomicAdd(W.page_ctr, 1u);
    pending = true;
  };
  auto put = [&](u32 idx, u32 pos, u32 h) {            // idx: rank inside the warp's run
    const u32 pg = (idx >> SS_PAGE_SHIFT) == cur_seq ? cur : nxt;
    W.pent[((u64)pg << SS_PAGE_SHIFT) | (idx & (SS_PAGE - 1))] = make_uint2(pos, h);
  };

  const u32 lt_mask = (1u << lane) - 1;
  u32 run_s = 0, run_c = 0;                            // height / #breaks since the start of the run
  bool sat = false;                                    // a cell beyond the reference's int16 range (2558-2573)
  int stg = 0;
  // chromosome of the current 8192-cell block; the next block's is fetched one block ahead
  const u32 last_blk = (s1 - 1) >> 4;
  int c = -1, c_next = L.blk2chrom[s0 >> 4];
  u64 off = 0;
  u32 len = 0;
  bool act = false;
  u32 sp = s0;
  while (sp < s1) {
    const u32 blk = sp >> 4;
    const int cn = c_next;
    if (blk < last_blk) c_next = L.blk2chrom[blk + 1];
    if (cn != c) {                                     // ~25 times per genome
      c = cn;
      off = L.off[c];
      len = L.len[c];
      act = (L.flags[c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
    }
    const u32 blk_end = min(s1, (blk + 1) << 4);
    u32 jb = (u32)(((u64)sp << 9) - off);              // chromosome position of the span's first cell
    if (jb == 0 && lane == 0) W.marks[c] = make_uint4(gw, run_s, run_c, 1u);
    for (; sp < blk_end; sp++, jb += 512, src += 128) {
      cp_async_wait<NSTAGE - 1>();
      __syncwarp();
      const int4* wst = wring + stg * 128 - lane;      // stage base (lane-independent)
      if (act && jb >= 1 && (u64)jb + 512 <= (u64)len) {
        // ~94 % of the cells are zero: list the non-zero 16-byte chunks, work only on them
        u32 nnz = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int4 x = wst[r * 32 + lane];
          const bool nz = (x.x | x.y | x.z | x.w) != 0;
          const u32 M = __ballot_sync(GR_FULL, nz);
          if (nz) wlist[nnz + __popc(M & lt_mask)] = (unsigned char)(r * 32 + lane);
          nnz += __popc(M);
        }
        if (lane < 16) wbm[lane] = 0;
        __syncwarp();
        for (u32 base = 0; base < nnz; base += 32) {
          resolve();
          const u32 n = base + lane;
          const bool on = n < nnz;
          int4 x = make_int4(0, 0, 0, 0);
          u32 q = 0;
          if (on) { q = wlist[n]; x = wst[q]; }
          const u32 m4 = (x.x != 0 ? 1u : 0u) | (x.y != 0 ? 2u : 0u) | (x.z != 0 ? 4u : 0u) | (x.w != 0 ? 8u : 0u);
          sat |= cell_saturated_dense(x.x) | cell_saturated_dense(x.y) | cell_saturated_dense(x.z) | cell_saturated_dense(x.w);
          const u32 c4 = __popc(m4);
          const u32 s1_ = (u32)x.x, s2_ = s1_ + (u32)x.y, s3_ = s2_ + (u32)x.z, s4_ = s3_ + (u32)x.w;
          const u32 inc_s = warp_incl_scan_u32(s4_, lane), inc_c = warp_incl_scan_u32(c4, lane);
          if (on) {
            const u32 h0 = run_s + inc_s - s4_;        // height before this chunk
            u32 idx = run_c + inc_c - c4;
            const u32 p0 = jb + q * 4;
            if (m4 & 1u) put(idx++, p0, h0);
            if (m4 & 2u) put(idx++, p0 + 1, h0 + s1_);
            if (m4 & 4u) put(idx++, p0 + 2, h0 + s2_);
            if (m4 & 8u) put(idx++, p0 + 3, h0 + s3_);
            atomicOr(&wbm[q >> 3], m4 << ((q & 7) * 4));
            // every break of an interior span is a non-zero cell and vice versa: clearing the
            // chunk leaves the delta array all zero for the next sample (no 4 B/bp memset)
            if (zero_after) const_cast<int4*>(src)[(int)q - lane] = make_int4(0, 0, 0, 0);
          }
          run_s += __shfl_sync(GR_FULL, inc_s, 31);
          run_c += __shfl_sync(GR_FULL, inc_c, 31);
          if ((run_c >> SS_PAGE_SHIFT) > cur_seq) advance();
        }
        __syncwarp();
        if (lane < 16) bitmap[(u64)sp * 16 + lane] = wbm[lane];
      } else {
        // dense: chromosome ends, inactive chromosomes.  16 consecutive cells per lane.
        int d[SC_ITEMS];
        sc_load_items(wst, lane, d);
        u32 run = 0, m = 0;
        const u32 j0 = jb + lane * SC_ITEMS;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
          run += (u32)d[i];
          sat |= cell_saturated_dense(d[i]);
          const u32 jj = j0 + i;
          const bool brk = (jj == len) || (d[i] != 0 && jj >= 1 && jj < len);
          m |= (brk ? 1u : 0u) << i;
        }
        if (!act) m = 0;
        const u32 cnt = __popc(m);
        const u32 wi_sum = warp_incl_scan_u32(run, lane), wi_cnt = warp_incl_scan_u32(cnt, lane);
        const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
        if (!(lane & 1)) bitmap[(u64)sp * 16 + (lane >> 1)] = m | (hi << 16);
        const u32 tot_c = __shfl_sync(GR_FULL, wi_cnt, 31);
        if (tot_c) {
          // up to 512 entries: page by page
          const u32 first_idx = run_c + (wi_cnt - cnt);
          const u32 last_seq = (run_c + tot_c - 1) >> SS_PAGE_SHIFT;
          for (;;) {
            resolve();
            u32 idx = first_idx, rr = run_s + (wi_sum - run);
#pragma unroll
            for (int i = 0; i < SC_ITEMS; i++) {
              if (m & (1u << i)) {
                if ((idx >> SS_PAGE_SHIFT) == cur_seq) put(idx, j0 + i, rr);
                idx++;
              }
              rr += (u32)d[i];
            }
            if (cur_seq >= last_seq) break;
            advance();
          }
          if (((run_c + tot_c) >> SS_PAGE_SHIFT) > cur_seq) advance();
        }
        run_s += __shfl_sync(GR_FULL, wi_sum, 31);
        run_c += tot_c;
        if (zero_after) {
#pragma unroll
          for (int i = 0; i < SC_ITEMS; i++)
            if (d[i] != 0) delta[(u64)sp * 512 + lane * SC_ITEMS + i] = 0;
        }
      }
      __syncwarp();                                    // every lane is done with the stage: refill it
      issue(sp + NSTAGE < s1, stg, src + (u64)NSTAGE * 128);
      stg = stg + 1 == NSTAGE ? 0 : stg + 1;
    }
  }
  resolve();                                           // a page still on order gets its (unused) label
  if (sat) atomicOr(err, GR_DE_SAT);
  if (lane == 0) W.warp_tot[gw] = make_uint2(run_s, run_c);
  cp_async_wait<0>();
}

// K2b: one CTA.  Exclusive scan of the per-warp totals; chromosome starts; tail check.
__global__ void __launch_bounds__(1024)
k_scan_fix(DevLayout L, StreamWs W, DevRle out, int* __restrict__ err, u32 nwarps) {
  __shared__ u32 sh_s[1024];
  __shared__ u64 sh_c[1024];
  const int t = threadIdx.x;
  const u32 per = (nwarps + 1023) / 1024;
  const u32 a = min(nwarps, t * per), b = min(nwarps, a + per);
  u32 s = 0;
  u64 c = 0;
  for (u32 i = a; i < b; i++) { const uint2 v = W.warp_tot[i]; s += v.x; c += v.y; }
  sh_s[t] = s; sh_c[t] = c;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    u32 vs = 0; u64 vc = 0;
    if (t >= o) { vs = sh_s[t - o]; vc = sh_c[t - o]; }
    __syncthreads();
    sh_s[t] += vs; sh_c[t] += vc;
    __syncthreads();
  }
  u32 bs = sh_s[t] - s;
  u64 bc = sh_c[t] - c;
  for (u32 i = a; i < b; i++) {
    const uint2 v = W.warp_tot[i];
    W.warp_base[i] = make_ulonglong2(bs, bc);
    bs += v.x; bc += v.y;
  }
  const u64 total = sh_c[1023];
  __syncthreads();
  for (int ch = t; ch < L.nchrom; ch += 1024) {
    if (L.off[ch] == ~0ull) continue;
    const uint4 mk = W.marks[ch];
    const ulonglong2 wb = W.warp_base[mk.x];
    out.chrom_start[ch] = wb.y + mk.z;
    if ((u32)wb.x + mk.y != 0) atomicOr(err, GR_DE_TAIL);   // previous chromosome did not return to 0 (2283-2289)
  }
  if (t == 0) { *out.total = total; out.chrom_start[L.nchrom] = total; }
  __syncthreads();
  if (t == 0) {
    u64 next = total;
    for (int ch = L.nchrom - 1; ch >= 0; ch--) {
      if (L.off[ch] == ~0ull) out.chrom_start[ch] = next;
      else next = out.chrom_start[ch];
    }
  }
}

// K2c: every page goes to its final rank; heights become the reference's floats.
// 16 B per break (1.5 GB per hg38 sample) in 0.37 ms: 4.1 TB/s with ONE page per half CTA and
// step; keeping four pages in flight (SP_UNROLL 4) was measured slower (0.41 ms), so the
// three-deep lookup chain (page -> owner -> base) is not what bounds it.
#define SP_UNROLL 1
// Tried and removed: taking the per-chromosome sums of len * val here (shared-memory u64 accumulators per CTA, the
// page-first entries by a second small kernel) instead of in k_rle_moment -- measured slower on the B200 (this
// kernel 0.42 -> 1.04 ms per step at two ranks for the 0.25 ms k_rle_moment took).
__global__ void __launch_bounds__(2 * SS_PAGE)
k_scan_place(StreamWs W, DevRle out, int* __restrict__ err, float excl_val) {
  __shared__ float4 sm_lut[120];
  units_lut_fill(sm_lut, threadIdx.x, 2 * SS_PAGE);
  __syncthreads();
  const u32 npages = min(*W.page_ctr, W.max_pages);
  const u32 t = threadIdx.x & (SS_PAGE - 1), half = threadIdx.x >> SS_PAGE_SHIFT;
  const u32 pstep = gridDim.x * 2;
  bool neg = false;
  for (u32 p0 = blockIdx.x * 2 + half; p0 < npages; p0 += pstep * SP_UNROLL) {
    uint2 meta[SP_UNROLL], e[SP_UNROLL];
    u32 tot[SP_UNROLL];
    ulonglong2 wb[SP_UNROLL];
#pragma unroll
    for (int k = 0; k < SP_UNROLL; k++) {
      const u32 p = min(p0 + k * pstep, npages - 1);
      meta[k] = W.page_meta[p];
      e[k] = W.pent[((u64)p << SS_PAGE_SHIFT) + t];    // may be stale past the page's fill: not used then
    }
#pragma unroll
    for (int k = 0; k < SP_UNROLL; k++) {
      tot[k] = W.warp_tot[meta[k].x].y;
      wb[k] = W.warp_base[meta[k].x];
    }
#pragma unroll
    for (int k = 0; k < SP_UNROLL; k++) {
      const u32 first = meta[k].y << SS_PAGE_SHIFT;
      if (p0 + k * pstep >= npages || first + t >= tot[k]) continue;   // beyond the owner's last entry / a page held in reserve
      const int N = (int)((u32)wb[k].x + e[k].y);
      neg |= N < 0;
      const u64 rank = wb[k].y + first + t;
      out.end[rank] = e[k].x & 0x7fffffffu;
      // bit 31: the interval lies in a -E region -- 0.0f in the experimental pileup (2248), SKIP in the control's (2124)
      out.val[rank] = (e[k].x >> 31) ? excl_val : units_to_val_lut(sm_lut, N < 0 ? 0 : N);
    }
  }
  if (neg) atomicOr(err, GR_DE_PILE);                  // ERRPILE 1921, 1969
}

// chromosomes without slots start where the next one does
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

size_t dense_scan_ws_bytes(u64 cap, int nchrom) {
  const u64 max_pages = (cap / SS_PAGE + SS_SPARE_PAGES + 3) & ~1ull;
  return (size_t)(max_pages * SS_PAGE * 8 + max_pages * 8 + SS_MAX_WARPS * (8 + 16) + (u64)nchrom * 16 + 256);
}

static int scan_stream_warps() {                       // warps of one K2a launch (2 CTAs of 16 warps per SM)
  static int n = 0;
  if (!n) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    n = sms * 2 * 16;
    if (n > SS_MAX_WARPS) n = SS_MAX_WARPS;
  }
  return n;
}

static StreamWs stream_ws(const ScanScratch& sc, int nchrom) {
  StreamWs W;
  const u64 max_pages = (sc.cap / SS_PAGE + SS_SPARE_PAGES + 3) & ~1ull;   // even: keeps the 16-byte arrays aligned
  char* p = (char*)sc.ws;
  W.pent = (uint2*)p; p += max_pages * SS_PAGE * 8;
  W.page_meta = (uint2*)p; p += max_pages * 8;
  W.warp_base = (ulonglong2*)p; p += SS_MAX_WARPS * 16;
  W.warp_tot = (uint2*)p; p += SS_MAX_WARPS * 8;
  W.marks = (uint4*)p; p += (u64)nchrom * 16;
  W.page_ctr = (u32*)p;
  W.max_pages = (u32)max_pages;
  return W;
}

template <int NSTAGE, int CPS>
static void launch_scan_stream_t(cudaStream_t s, const DevLayout& L, int32_t* delta, const StreamWs& W,
                                 u32* bitmap, int* err, int zero_after, int sms) {
  const size_t smem = (size_t)16 * NSTAGE * 2048;
  if (first_use_on_device(0 + NSTAGE)) cudaFuncSetAttribute(k_scan_stream<NSTAGE, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const u32 nwarps = (u32)scan_stream_warps();
  const u32 nspans = (u32)(L.T / 512);
  const u32 R = (nspans + nwarps - 1) / nwarps;
  k_scan_stream<NSTAGE, CPS><<<nwarps / 16, 512, smem, s>>>(delta, L, W, bitmap, err, nspans, R, zero_after);
  GR_NOTE_LAUNCH();
}

// K2a alone (the stage the roofline is quoted on) ...
void launch_dense_scan(cudaStream_t s, const DevLayout& L, int32_t* delta,
                       const ScanScratch& sc, u32* bitmap, int* err, int zero_after) {
  static int sms = 0, nstage = 0;
  if (!sms) {
    const char* e = getenv("GR_SCAN_STAGES");          // tuning knob; 2 and 3 measured equal
    nstage = e ? atoi(e) : 2;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const StreamWs W = stream_ws(sc, L.nchrom);
  cudaMemsetAsync(W.page_ctr, 0, 4, s);
  if (nstage == 3) launch_scan_stream_t<3, 2>(s, L, delta, W, bitmap, err, zero_after, sms);
  else launch_scan_stream_t<2, 2>(s, L, delta, W, bitmap, err, zero_after, sms);
}

// ... and K2b + K2c, which put the breaks where the rest of the pipeline expects them
void launch_scan_place(cudaStream_t s, const DevLayout& L, const ScanScratch& sc, DevRle out, int* err, u32 owners,
                       float excl_val) {
  const StreamWs W = stream_ws(sc, L.nchrom);
  k_scan_fix<<<1, 1024, 0, s>>>(L, W, out, err, owners ? owners : (u32)scan_stream_warps()); GR_NOTE_LAUNCH();
  k_scan_place<<<148 * 4, 2 * SS_PAGE, 0, s>>>(W, out, err, excl_val); GR_NOTE_LAUNCH();
}

// ============================================================================
// K1+K2 fused: the delta array never goes to HBM.
//
// The bucketed build above writes every 8192-cell block of the delta array once (4 B/cell)
// and the streaming scan reads it once (4 B/cell): 2 x 12.35 GB per hg38 sample for an array
// whose cells are ~97 % zero.  Here the block is assembled in shared memory and scanned where
// it lies: HBM sees the bucket entries (4 B per record), the breaks (8 B each) and the break
// bitmap (1 bit per cell).
//
// Buckets: one 4-byte EVENT entry per record in the block of its start; a record whose end
// lies in a later block gets a second, end-only entry in that block -- so every block is
// self-contained (no spill list, no limit on the interval length).
//   bits 0-12 cell offset inside the block | 13-25 interval length (kind 0) | 26-29 count |
//   30-31 kind: 0 = start and end in this block, 1 = start only, 2 = end only, 3 = -E region
//   boundary (no weight: the cell becomes a break whatever its delta, Genrich.c:2241)
// k_fb_scan: a CTA owns a contiguous run of blocks and carries (height, #breaks) relative to
// the start of its run, exactly like a warp of k_scan_stream does for its run of spans; the
// breaks go to the same pages and k_scan_fix / k_scan_place finish the job (owner = CTA).
// Per block: events -> smem cells (atomicAdd) + a 256-word occupancy bitmap (atomicOr); each
// thread then owns one bitmap word = 32 cells and only looks at the cells whose bit is set,
// clearing them behind itself, so the 32 KB of cells are zeroed once per kernel, not per block.
#define FB_KIND_BOTH 0u
#define FB_KIND_START 1u
#define FB_KIND_END 2u
#define FB_KIND_MARK 3u                                // -E region boundary: forces a break, carries no weight
__device__ __forceinline__ u32 fb_entry(u32 so, u32 span, int cnt, u32 kind) {
  return so | (span << 13) | ((u32)cnt << 26) | (kind << 30);
}

// Both passes keep FB_UNROLL records per thread in flight (one load / one cursor atomic at a
// time left them waiting on the memory latency: ncu long-scoreboard stalls 32 and 51 per issue).
#define FB_UNROLL 4
template <bool PACKED>
__global__ void __launch_bounds__(256)
k_fb_count(const void* __restrict__ recs, u64 n, DevLayout L, u32* __restrict__ blk_cnt,
           int* __restrict__ err, u64* __restrict__ clamped) {
  const u64 stride = (u64)gridDim.x * (256 * FB_UNROLL);
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i0 = (u64)blockIdx.x * (256 * FB_UNROLL) + threadIdx.x; i0 < n; i0 += stride) {
    int4 r[FB_UNROLL];
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++)
      if (i0 + k * 256 < n) r[k] = load_raw<PACKED>(recs, i0 + k * 256);
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++) {
      if (i0 + k * 256 >= n) break;
      u64 s_slot; u32 span; int w;
      if (!decode_raw<PACKED>(r[k], L, s_slot, span, w, e_local, c_local)) continue;
      const u64 bs = s_slot >> GR_BLOCK_SHIFT, be = (s_slot + span) >> GR_BLOCK_SHIFT;
      atomicAdd(blk_cnt + bs, 1u);
      if (be != bs) atomicAdd(blk_cnt + be, 1u);
    }
  }
  if (e_local) atomicOr(err, e_local);       // errors and clamp counts are reported by this pass only
  if (c_local && clamped) atomicAdd(clamped, (u64)c_local);
}

template <bool PACKED>
__global__ void __launch_bounds__(256)
k_fb_move(const void* __restrict__ recs, u64 n, DevLayout L, u32* __restrict__ cursor, u32* __restrict__ bucketed,
          const u32* __restrict__ sat_res, const u32* __restrict__ skip_bits, u64 seg_base) {
  // records the reference would have skipped (int16 saturation, k_sat_resolve): none in any ordinary sample
  const bool any_skip = sat_res && (sat_res[0] | sat_res[1]);
  const u32 omask = GR_BLOCK_SLOTS - 1;
  const u64 stride = (u64)gridDim.x * (256 * FB_UNROLL);
  int e_local = 0;
  u32 c_local = 0;
  for (u64 i0 = (u64)blockIdx.x * (256 * FB_UNROLL) + threadIdx.x; i0 < n; i0 += stride) {
    int4 r[FB_UNROLL];
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++)
      if (i0 + k * 256 < n) r[k] = load_raw<PACKED>(recs, i0 + k * 256);
    bool ok[FB_UNROLL], two[FB_UNROLL];
    u32 e0[FB_UNROLL], e1[FB_UNROLL], bs[FB_UNROLL], be[FB_UNROLL], p0[FB_UNROLL], p1[FB_UNROLL];
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++) {
      u64 s_slot = 0; u32 span = 0; int w = 120;
      ok[k] = i0 + k * 256 < n && decode_raw<PACKED>(r[k], L, s_slot, span, w, e_local, c_local);
      if (any_skip && ok[k]) {
        const u64 gi = seg_base + i0 + k * 256;
        ok[k] = !((skip_bits[gi >> 5] >> (gi & 31)) & 1u);
      }
      const u64 e_slot = s_slot + span;
      bs[k] = (u32)(s_slot >> GR_BLOCK_SHIFT);
      be[k] = (u32)(e_slot >> GR_BLOCK_SHIFT);
      two[k] = ok[k] && be[k] != bs[k];
      const u32 so = (u32)s_slot & omask;
      const int cnt = 120 / w;
      e0[k] = two[k] ? fb_entry(so, 0, cnt, FB_KIND_START) : fb_entry(so, span, cnt, FB_KIND_BOTH);
      e1[k] = fb_entry((u32)e_slot & omask, 0, cnt, FB_KIND_END);
    }
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++) {              // the cursor atomics of all records, back to back
      p0[k] = ok[k] ? atomicAdd(cursor + bs[k], 1u) : 0u;
      p1[k] = two[k] ? atomicAdd(cursor + be[k], 1u) : 0u;
    }
#pragma unroll
    for (int k = 0; k < FB_UNROLL; k++) {
      if (ok[k]) bucketed[p0[k]] = e0[k];
      if (two[k]) bucketed[p1[k]] = e1[k];
    }
  }
}

// Tried and removed: a two-level form (coarse bins sorted in shared memory, then one CTA per coarse bin: no
// per-entry global atomic).  Measured on the B200 at 2.64 ms per hg38 ChIP step against 2.11 ms for the count /
// move pair above -- the 8-byte intermediate items cost more HBM traffic than the atomics cost time.

// ---- the reference's int16 saturation rule (saveInterval, Genrich.c:2558-2573) -------------------
// The reference keeps (int16 cov, 8-bit frac) per delta cell and, IN ARRIVAL ORDER, drops an interval
// whole when the cell of its start already holds cov == INT16_MAX, else when the cell of its end holds
// cov == INT16_MIN.  A cell can get there only if 32767 or more interval starts (ends) land on it, so
// only a block with SAT_MIN_EVENTS entries can hold one: k_sb_scan1 raises a flag for such blocks, and
// this kernel -- one CTA, launched for every sample, returning at once unless the flag is up (no
// ordinary sample raises it: a hot spot of > 32 k identical fragment ends, e.g. chrM or an amplicon) --
//   1. lists the suspect blocks and sums, per cell of them, the weights of the starts and of the ends;
//   2. marks the HOT cells: those whose starts (ends) alone could take cov to the limit;
//   3. replays, in arrival order, the records that touch a hot cell, with the exact cell state (an
//      integer count of 1/120ths; cov is a function of it, sat_cov) of the hot cells only -- the decision
//      for a record depends on nothing else, and a cell that is not hot can never sit at a limit;
//   4. takes the dropped records out: a bit per record (k_fb_move leaves them out), the block counts
//      and their chunk sums corrected, a list of (arrival index, overflow | underflow) for the host's
//      warnings (gr_sample_skipped).
// sat_res (3 words, written in every launch): dropped for overflow, for underflow, list entries.
// integer part of the reference's cell for a count N of 1/120ths (canonical mixed-radix form, also for N < 0)
__device__ __forceinline__ int sat_cov(i64 N) {
  int r = (int)(N % 120);
  if (r < 0) r += 120;
  const int s = (2 * (r % 3)) % 3, t = (3 * (r % 5)) % 5, e = (4 * s + 4 * t - r) & 7;
  return (int)((N - (15 * e + 20 * s + 12 * t)) / 120);
}
#define SAT_HOT_STARTS (32767ll * 120)             /* cov == INT16_MAX needs N >= 32767 * 120 */
#define SAT_HOT_ENDS (32768ll * 120 - 193)         /* cov == INT16_MIN needs N <= -32768 * 120 + 193 (the fraction's maximum) */
__global__ void __launch_bounds__(1024)
k_sat_resolve(const u32* __restrict__ flag, const SatSeg* __restrict__ segs, int nseg, DevLayout L,
              u32* __restrict__ blk_cnt, u32* __restrict__ chunk_sum, u32 nblocks, ulonglong2* __restrict__ cells,
              u32* __restrict__ skip_bits, u64 nbits, u64* __restrict__ list, u32 list_cap, u32* __restrict__ sat_res,
              int* __restrict__ err) {
  __shared__ u32 sm_sb[SAT_MAX_BLOCKS];
  __shared__ u32 sm_nsb, sm_nhot, sm_pend_n;
  __shared__ u32 sm_wcnt[32];
  __shared__ struct { u64 gi; u32 s_id, e_id, bs, be; int w; } sm_pend[1024];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t < 3) sat_res[t] = 0;
  if (*flag == 0) return;
  if (t == 0) { sm_nsb = 0; sm_nhot = 0; }
  __syncthreads();
  for (u32 b = t; b < nblocks; b += 1024)
    if (blk_cnt[b] >= SAT_MIN_EVENTS) { const u32 k = atomicAdd(&sm_nsb, 1u); if (k < SAT_MAX_BLOCKS) sm_sb[k] = b; }
  __syncthreads();
  const u32 nsb = sm_nsb;
  if (nsb > SAT_MAX_BLOCKS) {                          // more hot spots than this path is sized for: reported, not guessed
    if (t == 0) atomicOr(err, GR_DE_SAT);
    return;
  }
  auto find = [&](u32 b) -> int {                      // index of block b among the suspect ones, -1 if it is none
    for (u32 k = 0; k < nsb; k++) if (sm_sb[k] == b) return (int)k;
    return -1;
  };
  for (u64 i = t; i < (u64)nsb * GR_BLOCK_SLOTS; i += 1024) cells[i] = make_ulonglong2(0, 0);
  for (u64 i = t; i < (nbits + 31) / 32; i += 1024) skip_bits[i] = 0;
  __syncthreads();
  // ---- 1: weights of the starts (.x) and of the ends (.y) per cell of the suspect blocks
  for (int g = 0; g < nseg; g++) {
    const SatSeg sg = segs[g];
    for (u64 i = t; i < sg.n; i += 1024) {
      u64 s_slot; u32 span; int w, e_local = 0; u32 c_local = 0;
      const bool ok = sg.packed ? decode_record<true>(sg.d, i, L, s_slot, span, w, e_local, c_local)
                                : decode_record<false>(sg.d, i, L, s_slot, span, w, e_local, c_local);
      if (!ok) continue;
      const u64 e_slot = s_slot + span;
      const int ks = find((u32)(s_slot >> GR_BLOCK_SHIFT)), ke = find((u32)(e_slot >> GR_BLOCK_SHIFT));
      if (ks >= 0) atomicAdd(&cells[(u64)ks * GR_BLOCK_SLOTS + (s_slot & (GR_BLOCK_SLOTS - 1))].x, (u64)w);
      if (ke >= 0) atomicAdd(&cells[(u64)ke * GR_BLOCK_SLOTS + (e_slot & (GR_BLOCK_SLOTS - 1))].y, (u64)w);
    }
  }
  __syncthreads();
  // ---- 2: hot cells.  .x becomes the flag, .y the running state (starts at 0: the array was zero)
  for (u64 i = t; i < (u64)nsb * GR_BLOCK_SLOTS; i += 1024) {
    const ulonglong2 v = cells[i];
    const bool hot = (i64)v.x >= SAT_HOT_STARTS || (i64)v.y >= SAT_HOT_ENDS;
    if (hot) atomicAdd(&sm_nhot, 1u);
    cells[i] = make_ulonglong2(hot ? 1ull : 0ull, 0ull);
  }
  __syncthreads();
  if (sm_nhot == 0) return;                            // a full block, but no cell of it can saturate
  // ---- 3 + 4: the records that touch a hot cell, 1024 at a time in arrival order; thread 0 decides
  u32 n_over = 0, n_under = 0, n_list = 0;             // thread 0's
  for (int g = 0; g < nseg; g++) {
    const SatSeg sg = segs[g];
    for (u64 i0 = 0; i0 < sg.n; i0 += 1024) {
      const u64 i = i0 + t;
      bool touch = false;
      u32 s_id = ~0u, e_id = ~0u, bs = 0, be = 0;
      int w = 0;
      if (i < sg.n) {
        u64 s_slot; u32 span; int e_local = 0; u32 c_local = 0;
        const bool ok = sg.packed ? decode_record<true>(sg.d, i, L, s_slot, span, w, e_local, c_local)
                                  : decode_record<false>(sg.d, i, L, s_slot, span, w, e_local, c_local);
        if (ok) {
          const u64 e_slot = s_slot + span;
          bs = (u32)(s_slot >> GR_BLOCK_SHIFT); be = (u32)(e_slot >> GR_BLOCK_SHIFT);
          const int ks = find(bs), ke = find(be);
          if (ks >= 0) { const u32 id = (u32)ks * GR_BLOCK_SLOTS + (u32)(s_slot & (GR_BLOCK_SLOTS - 1)); if (cells[id].x) s_id = id; }
          if (ke >= 0) { const u32 id = (u32)ke * GR_BLOCK_SLOTS + (u32)(e_slot & (GR_BLOCK_SLOTS - 1)); if (cells[id].x) e_id = id; }
          touch = s_id != ~0u || e_id != ~0u;
        }
      }
      const u32 bal = __ballot_sync(GR_FULL, touch);
      if (lane == 0) sm_wcnt[wid] = __popc(bal);
      __syncthreads();
      u32 pos = __popc(bal & ((1u << lane) - 1));
      for (int k = 0; k < wid; k++) pos += sm_wcnt[k];
      if (touch) { sm_pend[pos].gi = sg.base + i; sm_pend[pos].s_id = s_id; sm_pend[pos].e_id = e_id;
                   sm_pend[pos].bs = bs; sm_pend[pos].be = be; sm_pend[pos].w = w; }
      if (t == 1023) sm_pend_n = pos + (touch ? 1u : 0u);
      __syncthreads();
      if (t == 0) {
        const u32 np = sm_pend_n;
        for (u32 k = 0; k < np; k++) {
          const u32 a = sm_pend[k].s_id, b = sm_pend[k].e_id;
          const int ww = sm_pend[k].w;
          int kind = -1;
          if (a != ~0u && sat_cov((i64)cells[a].y) == 32767) kind = 0;            // 2558: overflow
          else if (b != ~0u && sat_cov((i64)cells[b].y) == -32768) kind = 1;      // 2566: underflow
          if (kind < 0) {
            if (a != ~0u) cells[a].y = (u64)((i64)cells[a].y + ww);              // 2576-2583
            if (b != ~0u) cells[b].y = (u64)((i64)cells[b].y - ww);
            continue;
          }
          const u64 gi = sm_pend[k].gi;
          skip_bits[gi >> 5] |= 1u << (gi & 31);
          if (kind) n_under++; else n_over++;
          if (n_list < list_cap) list[n_list++] = (gi << 1) | (u64)kind;
          const u32 xs = sm_pend[k].bs, xe = sm_pend[k].be;
          blk_cnt[xs] -= 1u; chunk_sum[xs / SB_CHUNK] -= 1u;           // the scan's second half follows
          if (xe != xs) { blk_cnt[xe] -= 1u; chunk_sum[xe / SB_CHUNK] -= 1u; }
        }
      }
      __syncthreads();
    }
  }
  if (t == 0) { sat_res[0] = n_over; sat_res[1] = n_under; sat_res[2] = n_list; }
}
void launch_sat_resolve(cudaStream_t s, const u32* flag, const void* segs, int nseg, const DevLayout& L, u32* blk_cnt,
                        u32* chunk_sum, void* cells, u32* skip_bits, u64 nbits, u64* list, u32 list_cap, u32* sat_res, int* err) {
  k_sat_resolve<<<1, 1024, 0, s>>>(flag, (const SatSeg*)segs, nseg, L, blk_cnt, chunk_sum, (u32)L.nblocks, (ulonglong2*)cells,
                                   skip_bits, nbits, list, list_cap, sat_res, err);
  GR_NOTE_LAUNCH();
}

// -E region boundaries (a few thousand at most): one pseudo entry each, so that the scan finds
// them in the occupancy bitmap like any other event.  cursor == NULL: count pass.
__global__ void k_fb_marks(const u64* __restrict__ marks, u32 n, u32* __restrict__ blk_cnt,
                           u32* __restrict__ cursor, u32* __restrict__ bucketed) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 slot = marks[i];
  const u32 b = (u32)(slot >> GR_BLOCK_SHIFT);
  if (!cursor) atomicAdd(blk_cnt + b, 1u);
  else bucketed[atomicAdd(cursor + b, 1u)] = fb_entry((u32)slot & (GR_BLOCK_SLOTS - 1), 0, 1, FB_KIND_MARK);
}
void launch_fb_marks(cudaStream_t s, const u64* marks, u32 n, u32* blk_cnt, u32* cursor, u32* bucketed) {
  if (!n) return;
  k_fb_marks<<<(n + 127) / 128, 128, 0, s>>>(marks, n, blk_cnt, cursor, bucketed); GR_NOTE_LAUNCH();
}

#define FB_WORDS (GR_BLOCK_SLOTS / 32)                 // occupancy / break bitmap words per block
// Which scan form runs is decided ON THE DEVICE, from the sample as it was bucketed: the rank form (a warp per
// block) is the fast one for ordinary blocks, but a block of thousands of entries -- a deep sample, or the
// pile-ups of an ATAC-seq peak -- is one warp's serial work there, and the CTA form takes it with 128 threads.
// Both kernels are launched; the one whose turn it is not returns at once (`when`: 0 run, 1 run if a quarter of
// the entries lie in blocks of >= FORM_HOT_MIN entries, 2 run if not).  stat: k_sb_scan1's words.
__device__ __forceinline__ bool form_skip(const u32* __restrict__ stat, const u32* __restrict__ blk_start, u32 nblocks, int when) {
  if (!when) return false;
  const bool hot = (u64)stat[1] * 4 > (u64)blk_start[nblocks];
  return when == 1 ? !hot : hot;
}
#define FB_RING 128                                    // page ring: sequence numbers in flight <= 2 * 33 + 2
// NT threads per CTA, each owning WPT = 256 / NT consecutive bitmap words (32 * WPT cells);
// PF entry registers per thread are fetched one block ahead (PF * NT = 512 entries).
template <int CPS, int NT, bool BED>
__global__ void __launch_bounds__(NT, CPS)
k_fb_scan(const u32* __restrict__ bucketed, const u32* __restrict__ blk_start, DevLayout L, StreamWs W,
          u32* __restrict__ bitmap, int* __restrict__ err, u32 nblocks, u32 R,
          const uint8_t* __restrict__ blk_bed /* NULL: no -E regions */,
          const u32* __restrict__ chrom_marks /* BED: region boundaries per chromosome, else NULL */,
          u32* __restrict__ chrom_ever /* BED: 1 once a sample of the context had reads on the chromosome */, int is_expt,
          const u32* __restrict__ stat, int when) {
  constexpr int WPT = FB_WORDS / NT, PF = 512 / NT, NW = NT / 32;
  if (form_skip(stat, blk_start, nblocks, when)) return;
  __shared__ int sm_cell[GR_BLOCK_SLOTS];
  __shared__ u32 sm_occ[FB_WORDS];
  __shared__ u32 sm_mark[BED ? FB_WORDS : 1];          // -E region boundaries of the block (rare)
  __shared__ u32 sm_pg[FB_RING];
  __shared__ u32 sm_ws[NW], sm_wc[NW];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const u32 owner = blockIdx.x;
  const u32 b0 = owner * R, b1 = min(b0 + R, nblocks);
  if (b0 >= b1) {
    if (t == 0) W.warp_tot[owner] = make_uint2(0, 0);
    return;
  }
  for (int i = t; i < GR_BLOCK_SLOTS; i += NT) sm_cell[i] = 0;
  for (int i = t; i < FB_WORDS; i += NT) { sm_occ[i] = 0; if (BED) sm_mark[i] = 0; }

  // bucket bounds of blocks b, b+1, b+2 (rolling; the entry for b+3 is fetched a block ahead)
  auto ld_start = [&](u32 i) { return blk_start[min(i, nblocks)]; };
  u32 sA = ld_start(b0), sB = ld_start(b0 + 1), sC = ld_start(b0 + 2);
  // upper bound of the breaks of a block: two cells per entry, plus the chromosome end
  auto ub_of = [&](u32 a, u32 b) { return min(2u * (b - a) + 1u, (u32)GR_BLOCK_SLOTS + 1u); };

  // pages (thread 0): have_seq = highest sequence number that has a page; a request for more
  // is issued one block before it is needed and its answer is only looked at a block later
  const u32 last_page = W.max_pages - 1;
  int have_seq = -1;
  u32 pend_p0 = 0;
  int pend_k = 0;
  auto page_take = [&]() {                               // label the pages of the answered request
    for (int i = 0; i < pend_k; i++) {
      u32 pg = pend_p0 + (u32)i;
      if (pg > last_page) { atomicOr(err, GR_DE_TABLE); pg = last_page; }
      have_seq++;
      W.page_meta[pg] = make_uint2(owner, (u32)have_seq);
      sm_pg[have_seq & (FB_RING - 1)] = pg;
    }
    pend_k = 0;
  };
  auto page_ask = [&](u32 upto_idx) {                    // make sure sequence upto_idx >> 8 will have a page
    const int target = (int)(upto_idx >> SS_PAGE_SHIFT);
    if (target > have_seq) {
      pend_k = target - have_seq;
      pend_p0 = atomicAdd(W.page_ctr, (u32)pend_k);
    }
  };
  if (t == 0) page_ask(ub_of(sA, sB));

  // the first PF rounds of a block's entries are fetched while the previous block is worked on
  u32 v[PF];
#pragma unroll
  for (int k = 0; k < PF; k++) {
    v[k] = 0;
    if (sA + k * NT + t < sB) v[k] = __ldcs(bucketed + sA + k * NT + t);
  }

  // A chromosome that has not had a single read so far -- in this EXPERIMENTAL sample or in any sample before it,
  // control samples included -- is one interval (len, 0.0f) in the reference, whatever -E regions lie on it
  // (savePileupExpt 2178-2182: its diff array, which all samples share, was never allocated): its region boundaries
  // are then not breaks.  Once a sample had reads there the array exists, and a later sample without reads on the
  // chromosome is cut at the region boundaries like any other.  (The control side of a read-less chromosome is
  // saveLambda's, regions included: the same intervals either way.)
  bool c_plain = false;
  auto apply = [&](u32 e) {
    const u32 so = e & (GR_BLOCK_SLOTS - 1), kind = e >> 30;
    const int w = 120 / (int)((e >> 26) & 15u);
    if (BED && kind == FB_KIND_MARK && c_plain) return;
    atomicOr(sm_occ + (so >> 5), 1u << (so & 31));
    if (BED && kind == FB_KIND_MARK) { atomicOr(sm_mark + (so >> 5), 1u << (so & 31)); return; }
    atomicAdd(sm_cell + so, kind == FB_KIND_END ? -w : w);
    if (kind == FB_KIND_BOTH) {
      const u32 eo = so + ((e >> 13) & (GR_BLOCK_SLOTS - 1));
      atomicAdd(sm_cell + eo, -w);
      atomicOr(sm_occ + (eo >> 5), 1u << (eo & 31));
    }
  };

  u32 run_s = 0, run_c = 0;                            // height / #breaks since the start of the run
  bool sat = false;                                    // a cell beyond the reference's int16 range (2558-2573)
  int c = -1;
  u32 c_last_blk = 0;                                  // last block of chromosome c
  u64 off = 0;
  u32 len = 0;
  bool act = false;
  __syncthreads();
  for (u32 b = b0; b < b1; b++) {
    if (c < 0 || b > c_last_blk) {                     // ~25 times per genome
      c = L.blk2chrom[b];
      off = L.off[c];
      len = L.len[c];
      c_last_blk = (u32)((off + len) >> GR_BLOCK_SHIFT);
      act = (L.flags[c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
      if (BED && chrom_marks) {
        // nothing but its own region boundaries in the chromosome's buckets?
        const bool none = blk_start[c_last_blk + 1] - blk_start[(u32)(off >> GR_BLOCK_SHIFT)] == chrom_marks[c];
        c_plain = is_expt && none && !chrom_ever[c];
        // (another CTA may read the word for the same chromosome later in this launch: it is only written when
        // `none` is false, and then c_plain is false whatever it holds)
        if (!none && t == 0) chrom_ever[c] = 1u;
      }
    }
    const u32 sD = ld_start(b + 3);                    // used two blocks from now
    const u32 jb = (u32)(((u64)b << GR_BLOCK_SHIFT) - off);       // chromosome position of the block's first cell
    if (t == 0) {
      if (jb == 0) W.marks[c] = make_uint4(owner, run_s, run_c, 1u);
      page_take();                                     // covers this block (asked for a block ago)
      page_ask(run_c + ub_of(sA, sB) + (b + 1 < b1 ? ub_of(sB, sC) : 0u));
    }
    const bool has_end = act && b == c_last_blk;       // cell `len` lies in this block
    u32* const bm_out = bitmap + (u64)b * FB_WORDS + t * WPT;
    if (sA == sB && !has_end) {                        // nothing in this block
      if (WPT == 1) bm_out[0] = 0;
      else if (WPT == 2) *reinterpret_cast<uint2*>(bm_out) = make_uint2(0, 0);
      else *reinterpret_cast<uint4*>(bm_out) = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < PF; k++)
        if (sB + k * NT + t < sC) v[k] = __ldcs(bucketed + sB + k * NT + t);
      sA = sB; sB = sC; sC = sD;
      continue;
    }
    // ---- events -> cells
#pragma unroll
    for (int k = 0; k < PF; k++)
      if (sA + k * NT + t < sB) apply(v[k]);
    for (u32 i = sA + PF * NT + t; i < sB; i += NT) apply(__ldcs(bucketed + i));
#pragma unroll
    for (int k = 0; k < PF; k++)                       // next block's entries: in flight during the scan
      if (sB + k * NT + t < sC) v[k] = __ldcs(bucketed + sB + k * NT + t);
    __syncthreads();
    // ---- this thread's 32 * WPT cells: sum of the deltas, break masks
    const u32 end_cell = len - jb;                     // meaningful if has_end
    u32 mo[WPT], m[WPT], mk[WPT], xm[WPT];             // occupied, breaks, region boundaries, "interval ending here is excluded"
    u32 s = 0, cnt = 0;
    const int cbase = t * (32 * WPT);
    const u32 jt = jb + (u32)cbase;
    // -E: bit 0 of the block's byte = the interval running into the block is inside a region,
    // bit 1 = the block holds region boundaries (then, and only then, sm_mark is looked at)
    const u32 bed = (BED && !c_plain) ? (u32)blk_bed[b] : 0u;
    u32 excl_in = bed & 1u;                            // state of the interval ending at this thread's first cell
    if (bed & 2u)
      for (int i = 0; i < t * WPT; i++) excl_in ^= __popc(sm_mark[i]) & 1u;
#pragma unroll
    for (int q = 0; q < WPT; q++) {
      mo[q] = sm_occ[t * WPT + q];
      mk[q] = (bed & 2u) ? sm_mark[t * WPT + q] : 0u;
      if (has_end && (int)(end_cell >> 5) == t * WPT + q) mo[q] |= 1u << (end_cell & 31);
      m[q] = 0; xm[q] = 0;
      for (u32 mm = mo[q]; mm; mm &= mm - 1) {
        const int bit = __ffs(mm) - 1;
        const int d = sm_cell[cbase + q * 32 + bit];
        const u32 j = jt + (u32)(q * 32 + bit);
        s += (u32)d;
        sat |= cell_saturated(d);
        const bool mark = (mk[q] >> bit) & 1u;
        const bool brk = (j == len) || (j >= 1u && j < len && (mark || (!excl_in && d != 0)));
        m[q] |= (brk ? 1u : 0u) << bit;
        xm[q] |= excl_in << bit;
        excl_in ^= mark ? 1u : 0u;                     // the region state flips AFTER the interval is closed (2256-2263)
      }
      if (!act) m[q] = 0;
      cnt += __popc(m[q]);
    }
    if (bed & 2u) {                                    // every thread has read what it needs of sm_mark
      __syncthreads();
#pragma unroll
      for (int q = 0; q < WPT; q++) sm_mark[t * WPT + q] = 0;
    }
#pragma unroll
    for (int q = 0; q < WPT; q++) sm_occ[t * WPT + q] = 0;
    const u32 wi_s = warp_incl_scan_u32(s, lane), wi_c = warp_incl_scan_u32(cnt, lane);
    if (lane == 31) { sm_ws[wid] = wi_s; sm_wc[wid] = wi_c; }
    __syncthreads();
    u32 h = run_s + wi_s - s, idx = run_c + wi_c - cnt, tot_s = 0, tot_c = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
      const u32 a = sm_ws[k], q = sm_wc[k];
      if (k < wid) { h += a; idx += q; }
      tot_s += a; tot_c += q;
    }
    // ---- emit, clearing the cells behind.  Bit 31 of the coordinate: the interval is excluded.
#pragma unroll
    for (int q = 0; q < WPT; q++) {
      for (u32 mm = mo[q]; mm; mm &= mm - 1) {
        const int bit = __ffs(mm) - 1;
        const int d = sm_cell[cbase + q * 32 + bit];
        sm_cell[cbase + q * 32 + bit] = 0;
        if ((m[q] >> bit) & 1u) {
          const u32 pg = sm_pg[(idx >> SS_PAGE_SHIFT) & (FB_RING - 1)];
          W.pent[((u64)pg << SS_PAGE_SHIFT) | (idx & (SS_PAGE - 1))] =
              make_uint2((jt + (u32)(q * 32 + bit)) | (((xm[q] >> bit) & 1u) << 31), h);
          idx++;
        }
        h += (u32)d;
      }
    }
    if (WPT == 1) bm_out[0] = m[0];
    else if (WPT == 2) *reinterpret_cast<uint2*>(bm_out) = make_uint2(m[0], m[WPT > 1 ? 1 : 0]);
    else *reinterpret_cast<uint4*>(bm_out) = make_uint4(m[0], m[WPT > 1 ? 1 : 0], m[WPT > 2 ? 2 : 0], m[WPT > 3 ? 3 : 0]);
    run_s += tot_s;
    run_c += tot_c;
    sA = sB; sB = sC; sC = sD;
    __syncthreads();                                   // cells, occupancy words and scan scratch are free again
  }
  if (sat) atomicOr(err, GR_DE_SAT);
  if (t == 0) {
    page_take();                                       // every page handed out carries a label
    W.warp_tot[owner] = make_uint2(run_s, run_c);
  }
}

// Tried and removed: k_fd_scan, a cell-array form with the entry stream staged by bulk copies (cp.async.bulk +
// mbarrier, two 8 KB tiles) and 32 cells per thread.  Measured on the B200 per launch: hg38 ChIP 3.6 ms (rank form
// 0.71), ATAC 4.3 ms (CTA form 2.8), 10 Gbp shard 3.2 ms (CTA form 4.1) -- a 2 % shorter step on one workload.

// Rank form -- the scan of samples with ordinary blocks (form_skip): warp-owned 8192-cell blocks WITHOUT a
// cell array.  A block of the hg38 workload holds ~200 entries = ~400 distinct event cells out of 8192;
// k_fb_scan spends its time in three CTA barriers per block and in a walk whose length is the fullest
// thread's (ncu: issue slots half empty).  Measured on the B200 (hg38, 50 M records): 0.71 ms per launch
// against 1.69 ms, same bits.
// Here a warp keeps only the block's 256-word occupancy bitmap and, per DISTINCT event cell, a
// sum and a position:
//   P1  entries -> occupancy bits
//   P2  exclusive popcount prefix per word              (rank of a cell = prefix + bits below it)
//   P3  entries again -> sum[rank] += +-w, pos[rank] = cell
//   P4  the dense list, 32 ranks per round: height = inclusive warp scan, break test, ballot ->
//       page entries; a cell whose deltas cancel (or that lies outside [1, len]) leaves the bitmap
//   P5  the occupancy words ARE the break bitmap: written out, cleared
// Every lane has work in every round, nothing is walked, no __syncthreads.  A block with more
// than CAP distinct cells takes several rounds of P3/P4, cut at word boundaries (so that the bits
// P4 clears never sit below a cell that still has to be ranked).  Output contract = k_fb_scan's
// with owner = warp (pages, warp_tot, marks), so k_scan_fix / k_scan_place follow unchanged.
// 120 / count without the division subroutine (one per entry otherwise): byte `count` of a 16-entry table
__device__ __forceinline__ int fr_weight(u32 count) {
  const u64 t = (count & 8u) ? 0x0808090A0A0C0D0Full : 0x1114181E283C7800ull;   // 120 / 8..15 | 120 / 0..7 (0 -> 0)
  return (int)((t >> ((count & 7u) * 8u)) & 0xffu);
}
#define FR_RING 128                                    // page ring per warp: <= 2 * 33 + 2 sequence numbers in flight
// PF: entry registers per lane (32 * PF entries of a block are prefetched while the previous block is worked on)
template <int CAP, int CPS, int PF>
__global__ void __launch_bounds__(128, CPS)
k_fr_scan(const u32* __restrict__ bucketed, const u32* __restrict__ blk_start, DevLayout L, StreamWs W,
          u32* __restrict__ bitmap, int* __restrict__ err, u32 nblocks, u32 R, const u32* __restrict__ stat, int when) {
  if (form_skip(stat, blk_start, nblocks, when)) return;
  __shared__ __align__(16) u32 sm_occ_all[4 * FB_WORDS];
  __shared__ __align__(16) u32 sm_pre_all[4 * FB_WORDS];
  __shared__ int sm_sum_all[4 * CAP];
  __shared__ unsigned short sm_pos_all[4 * CAP];
  __shared__ u32 sm_pg_all[4 * FR_RING];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u32* const sm_occ = sm_occ_all + wid * FB_WORDS;
  u32* const sm_pre = sm_pre_all + wid * FB_WORDS;
  int* const sm_sum = sm_sum_all + wid * CAP;
  unsigned short* const sm_pos = sm_pos_all + wid * CAP;
  u32* const sm_pg = sm_pg_all + wid * FR_RING;
  const u32 owner = blockIdx.x * 4 + wid;
  const u32 b0 = owner * R, b1 = min(b0 + R, nblocks);
  if (b0 >= b1) {
    if (lane == 0 && owner < SS_MAX_WARPS) W.warp_tot[owner] = make_uint2(0, 0);
    return;
  }
  for (int i = lane; i < FB_WORDS; i += 32) sm_occ[i] = 0;
  for (int i = lane; i < CAP; i += 32) sm_sum[i] = 0;

  // entries of block i: n_of(i) of them, starting at lo_of(i)
  auto n_of = [&](u32 i) -> u32 {
    if (i >= nblocks) return 0u;
    return blk_start[i + 1] - blk_start[i];
  };
  auto lo_of = [&](u32 i) -> u64 {
    return (u64)blk_start[min(i, nblocks)];
  };
  auto ub_of = [&](u32 n) { return min(2u * n + 1u, (u32)GR_BLOCK_SLOTS + 1u); };   // breaks of a block, upper bound
  u32 nA = n_of(b0), nB = n_of(b0 + 1);
  const u32* eA = bucketed + lo_of(b0);

  const u32 last_page = W.max_pages - 1;
  int have_seq = -1;
  u32 pend_p0 = 0;
  int pend_k = 0;
  auto page_take = [&]() {
    for (int i = 0; i < pend_k; i++) {
      u32 pg = pend_p0 + (u32)i;
      if (pg > last_page) { atomicOr(err, GR_DE_TABLE); pg = last_page; }
      have_seq++;
      W.page_meta[pg] = make_uint2(owner, (u32)have_seq);
      sm_pg[have_seq & (FR_RING - 1)] = pg;
    }
    pend_k = 0;
  };
  auto page_ask = [&](u32 upto_idx) {
    const int target = (int)(upto_idx >> SS_PAGE_SHIFT);
    if (target > have_seq) {
      pend_k = target - have_seq;
      pend_p0 = atomicAdd(W.page_ctr, (u32)pend_k);
    }
  };
  if (lane == 0) page_ask(ub_of(nA));

  u32 v[PF];
#pragma unroll
  for (int k = 0; k < PF; k++) {
    v[k] = 0;
    if ((u32)(k * 32 + lane) < nA) v[k] = __ldcs(eA + k * 32 + lane);
  }

  u32 run_s = 0, run_c = 0;                            // height / #breaks since the start of the run
  bool sat = false;
  int c = -1;
  u32 c_last_blk = 0;
  u64 off = 0;
  u32 len = 0;
  bool act = false;
  __syncwarp();
  for (u32 b = b0; b < b1; b++) {
    if (c < 0 || b > c_last_blk) {                     // ~25 times per genome
      c = L.blk2chrom[b];
      off = L.off[c];
      len = L.len[c];
      c_last_blk = (u32)((off + len) >> GR_BLOCK_SHIFT);
      act = (L.flags[c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
    }
    const u32 nC = n_of(b + 2);                        // used from the next block on
    const u32* const eB = bucketed + lo_of(b + 1);
    const u32 jb = (u32)(((u64)b << GR_BLOCK_SHIFT) - off);       // chromosome position of the block's first cell
    if (lane == 0) {
      if (jb == 0) W.marks[c] = make_uint4(owner, run_s, run_c, 1u);
      page_take();                                     // covers this block (asked for a block ago)
      page_ask(run_c + ub_of(nA) + (b + 1 < b1 ? ub_of(nB) : 0u));
    }
    const bool has_end = act && b == c_last_blk;       // cell `len` lies in this block
    uint4* const bm_out = reinterpret_cast<uint4*>(bitmap + (u64)b * FB_WORDS + lane * 8);
    if (nA == 0 && !has_end) {                         // nothing in this block
      bm_out[0] = make_uint4(0, 0, 0, 0);
      bm_out[1] = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < PF; k++)
        if ((u32)(k * 32 + lane) < nB) v[k] = __ldcs(eB + k * 32 + lane);
      nA = nB; nB = nC; eA = eB;
      continue;
    }
    const u32 end_cell = len - jb;                     // meaningful if has_end
    // ---- P1: occupancy
    auto mark = [&](u32 e) {
      const u32 so = e & (GR_BLOCK_SLOTS - 1);
      atomicOr(sm_occ + (so >> 5), 1u << (so & 31));
      if ((e >> 30) == FB_KIND_BOTH) {
        const u32 eo = so + ((e >> 13) & (GR_BLOCK_SLOTS - 1));
        atomicOr(sm_occ + (eo >> 5), 1u << (eo & 31));
      }
    };
#pragma unroll
    for (int k = 0; k < PF; k++)
      if ((u32)(k * 32 + lane) < nA) mark(v[k]);
    for (u32 i = PF * 32 + lane; i < nA; i += 32) mark(__ldg(eA + i));
    if (has_end && lane == 0) atomicOr(sm_occ + (end_cell >> 5), 1u << (end_cell & 31));
    __syncwarp();
    // ---- P2: exclusive prefix of the word popcounts (lane: words 8 * lane .. 8 * lane + 7)
    u32 n_pos;
    {
      const uint4 o0 = *reinterpret_cast<const uint4*>(sm_occ + lane * 8);
      const uint4 o1 = *reinterpret_cast<const uint4*>(sm_occ + lane * 8 + 4);
      const u32 c0 = __popc(o0.x), c1 = __popc(o0.y), c2 = __popc(o0.z), c3 = __popc(o0.w);
      const u32 c4 = __popc(o1.x), c5 = __popc(o1.y), c6 = __popc(o1.z), c7 = __popc(o1.w);
      const u32 mine = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
      const u32 inc = warp_incl_scan_u32(mine, lane);
      uint4 p0, p1;
      p0.x = inc - mine; p0.y = p0.x + c0; p0.z = p0.y + c1; p0.w = p0.z + c2;
      p1.x = p0.w + c3; p1.y = p1.x + c4; p1.z = p1.y + c5; p1.w = p1.z + c6;
      *reinterpret_cast<uint4*>(sm_pre + lane * 8) = p0;
      *reinterpret_cast<uint4*>(sm_pre + lane * 8 + 4) = p1;
      n_pos = __shfl_sync(GR_FULL, inc, 31);
    }
    __syncwarp();
    // ---- rounds of at most CAP distinct cells (one round unless the block is unusually full)
    u32 lo = 0;
    while (true) {
      u32 hi = n_pos;
      if (n_pos - lo > (u32)CAP) {                     // cut at a word boundary: words [0, w_hi) hold <= lo + CAP cells
        u32 k = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const int w = lane * 8 + q;
          k += (sm_pre[w] + __popc(sm_occ[w]) <= lo + (u32)CAP) ? 1u : 0u;
        }
        const u32 w_hi = __reduce_add_sync(GR_FULL, k);
        hi = w_hi < (u32)FB_WORDS ? sm_pre[w_hi] : n_pos;
      }
      const u32 span = hi - lo;
      const bool last_round = hi >= n_pos;
      // ---- P3: sums and positions by rank
      auto touch = [&](u32 so, int w) {
        const u32 wd = so >> 5;
        const u32 r = sm_pre[wd] + __popc(sm_occ[wd] & ((1u << (so & 31)) - 1u)) - lo;
        if (r < span) {
          if (w) atomicAdd(sm_sum + r, w);
          sm_pos[r] = (unsigned short)so;
        }
      };
      auto add = [&](u32 e) {
        const u32 so = e & (GR_BLOCK_SLOTS - 1), kind = e >> 30;
        const int w = fr_weight((e >> 26) & 15u);
        touch(so, kind == FB_KIND_END ? -w : w);
        if (kind == FB_KIND_BOTH) touch(so + ((e >> 13) & (GR_BLOCK_SLOTS - 1)), -w);
      };
#pragma unroll
      for (int k = 0; k < PF; k++)
        if ((u32)(k * 32 + lane) < nA) add(v[k]);
      for (u32 i = PF * 32 + lane; i < nA; i += 32) add(__ldg(eA + i));
      if (has_end && lane == 0) touch(end_cell, 0);
      if (last_round) {                                // next block's entries: in flight during P4 / P5
#pragma unroll
        for (int k = 0; k < PF; k++)
          if ((u32)(k * 32 + lane) < nB) v[k] = __ldcs(eB + k * 32 + lane);
      }
      __syncwarp();
      // ---- P4: heights, breaks
      for (u32 r0 = 0; r0 < span; r0 += 32) {
        const u32 r = r0 + lane;
        const bool on = r < span;
        int d = 0;
        u32 p = 0;
        if (on) { d = sm_sum[r]; sm_sum[r] = 0; p = sm_pos[r]; }
        const u32 inc = warp_incl_scan_u32((u32)d, lane);
        const u32 j = jb + p;
        sat |= cell_saturated(d);
        const bool brk = on && act && ((j == len) || (d != 0 && j >= 1u && j < len));
        const u32 bal = __ballot_sync(GR_FULL, brk);
        if (brk) {
          const u32 idx = run_c + __popc(bal & ((1u << lane) - 1u));
          const u32 pg = sm_pg[(idx >> SS_PAGE_SHIFT) & (FR_RING - 1)];
          W.pent[((u64)pg << SS_PAGE_SHIFT) | (idx & (SS_PAGE - 1))] = make_uint2(j, run_s + inc - (u32)d);
        } else if (on && act) {
          atomicAnd(sm_occ + (p >> 5), ~(1u << (p & 31)));        // not a break: leaves the bitmap
        }
        run_s += __shfl_sync(GR_FULL, inc, 31);
        run_c += __popc(bal);
      }
      __syncwarp();
      if (last_round) break;
      lo = hi;
    }
    // ---- P5: what is left of the occupancy words is the break bitmap
    {
      uint4 o0 = *reinterpret_cast<const uint4*>(sm_occ + lane * 8);
      uint4 o1 = *reinterpret_cast<const uint4*>(sm_occ + lane * 8 + 4);
      *reinterpret_cast<uint4*>(sm_occ + lane * 8) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sm_occ + lane * 8 + 4) = make_uint4(0, 0, 0, 0);
      if (!act) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }
      bm_out[0] = o0;
      bm_out[1] = o1;
    }
    nA = nB; nB = nC; eA = eB;
    __syncwarp();                                      // occupancy words are free again
  }
  if (sat) atomicOr(err, GR_DE_SAT);
  if (lane == 0) {
    page_take();
    W.warp_tot[owner] = make_uint2(run_s, run_c);
  }
}

void launch_fb_count(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                     u32* blk_cnt, int* err, u64* clamped) {
  if (!n) return;
  u64 blocks = (n + 256 * FB_UNROLL - 1) / (256 * FB_UNROLL);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (packed) k_fb_count<true><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, blk_cnt, err, clamped);
  else k_fb_count<false><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, blk_cnt, err, clamped);
  GR_NOTE_LAUNCH();
}
void launch_fb_move(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed, u32* cursor, u32* bucketed,
                    const u32* sat_res, const u32* skip_bits, u64 seg_base) {
  if (!n) return;
  u64 blocks = (n + 256 * FB_UNROLL - 1) / (256 * FB_UNROLL);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (packed) k_fb_move<true><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, cursor, bucketed, sat_res, skip_bits, seg_base);
  else k_fb_move<false><<<(unsigned)blocks, 256, 0, s>>>(recs, n, L, cursor, bucketed, sat_res, skip_bits, seg_base);
  GR_NOTE_LAUNCH();
}

static int fb_env(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// bucketed events -> breaks (pages) + break bitmap; launch_scan_place(..., owners) follows.
// k_fr_scan (warp-owned blocks, rank form; 9 CTAs x 4 warps per SM, 512 distinct cells per round) and k_fb_scan
// (CTA-owned blocks, cell array in shared memory) are both launched and the sample picks one (form_skip):
// measured on the B200, an hg38 ChIP sample (200 entries per block) takes 0.71 ms in the rank form and 1.69 ms in
// the CTA form; the 10 Gbp / 1 B fragment shard (2200 entries per block) 9.5 ms against 4.1 ms, an ATAC sample
// with 13 k cut sites on a block 3.75 against 2.75.  Contexts with -E regions take the CTA form (the region
// boundaries are weightless mark entries only it understands).  GR_FUSED_CTA=1 / GR_FUSED_RANK=1 force a form
// (tests, and the comparison the bench quotes); stat: k_sb_scan1's statistic words.
u32 launch_fb_scan(cudaStream_t s, const DevLayout& L, const u32* bucketed, const u32* blk_start,
                   const ScanScratch& sc, u32* bitmap, int* err, const uint8_t* blk_bed, const u32* chrom_marks,
                   u32* chrom_ever, int is_expt, const u32* stat) {
  const StreamWs W = stream_ws(sc, L.nchrom);
  cudaMemsetAsync(W.page_ctr, 0, 4, s);
  cudaMemsetAsync(W.warp_tot, 0, (size_t)SS_MAX_WARPS * sizeof(uint2), s);   // owners of the form that does not run: empty
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const u32 nb = (u32)L.nblocks;
  const int force_cta = blk_bed || fb_env("GR_FUSED_CTA", 0), force_rank = !force_cta && fb_env("GR_FUSED_RANK", 0);   // read per call: the tests switch it inside one process
  u32 owners = 0;
  if (!force_rank) {
    const u32 o = (u32)(sms * 6);
    const u32 R = (nb + o - 1) / o;
    const int when = force_cta ? 0 : 1;
    if (blk_bed) k_fb_scan<6, 128, true><<<o, 128, 0, s>>>(bucketed, blk_start, L, W, bitmap, err, nb, R, blk_bed, chrom_marks, chrom_ever, is_expt, stat, 0);
    else k_fb_scan<6, 128, false><<<o, 128, 0, s>>>(bucketed, blk_start, L, W, bitmap, err, nb, R, nullptr, nullptr, nullptr, 0, stat, when);
    GR_NOTE_LAUNCH();
    owners = o;
  }
  if (!force_cta) {
    u32 o = (u32)(sms * 9) * 4;
    if (o > SS_MAX_WARPS) o = SS_MAX_WARPS & ~3u;
    const u32 R = (nb + o - 1) / o;
    k_fr_scan<512, 9, 8><<<o / 4, 128, 0, s>>>(bucketed, blk_start, L, W, bitmap, err, nb, R, stat, force_rank ? 0 : 2);
    GR_NOTE_LAUNCH();
    if (o > owners) owners = o;
  }
  return owners;
}

// ============================================================================
// K2b: per chromosome, sum over its RLE intervals of (float)(end-start)*val
// (the float product of Genrich.c:2246 / 2018, accumulated there in a double).
// Here every float product is added EXACTLY in fixed point (integer part and
// 2^-40 fraction in separate u64 counters), so the result does not depend on the
// order of the atomics; the host rounds int + frac*2^-40 to a double once.
__global__ void __launch_bounds__(256)
k_rle_moment(DevRle r, int nchrom, u64* __restrict__ acc_int, u64* __restrict__ acc_frac) {
  __shared__ u64 sm_i[8], sm_f[8];
  const u64 n = *r.total;
  // each CTA owns one contiguous slice of the interval array, so its running
  // chromosome changes at most a handful of times: sums stay in registers and
  // reach the per-chromosome counters with O(#CTAs) atomics instead of O(n/256).
  // The chromosome of the slice and the index where it ends are block-uniform REGISTERS:
  // a round that stays below that index needs no search, no shared memory and no barrier
  // (the first version looked the round's chromosomes up through thread 0 and two barriers
  // per 1024 intervals: ~3 us of dependent L2 loads per round, 0.28 ms per hg38 sample).
  const u64 per = ((n + gridDim.x - 1) / gridDim.x + 1023) / 1024 * 1024;
  const u64 lo = (u64)blockIdx.x * per;
  const u64 hi = min(lo + per, n);
  if (lo >= hi) return;
  u64 pi = 0, pf = 0;
  int cur = chrom_of_index(r.chrom_start, nchrom, lo);  // chromosome the register sums belong to
  u64 cs = r.chrom_start[cur], nb = r.chrom_start[cur + 1];
  auto flush = [&]() {
    // block-wide: add (pi, pf) of all threads into chromosome `cur`
    u64 a = warp_sum_u64(pi), b = warp_sum_u64(pf);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sm_i[w] = a; sm_f[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 ti = 0, tf = 0;
      for (int k = 0; k < 8; k++) { ti += sm_i[k]; tf += sm_f[k]; }
      ti += tf >> 40;
      tf &= (1ull << 40) - 1;
      if (ti) atomicAdd(acc_int + cur, ti);
      if (tf) atomicAdd(acc_frac + cur, tf);
    }
    pi = 0; pf = 0;
  };
  for (u64 base = lo; base < hi; base += 1024) {       // 4 intervals per thread per round
    const u64 last = min(base + 1024, hi) - 1;
    if (last < nb) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const u64 i = base + j * 256 + threadIdx.x;
        if (i < hi) {
          const u32 e = r.end[i];
          const u32 st = (i == cs) ? 0u : r.end[i - 1];
          const float v = r.val[i];
          const float p = v < 0.0f ? 0.0f : __fmul_rn(__uint2float_rn(e - st), v);     // SKIP (-E region): not counted (2016)
          const u64 ip = (u64)p;                       // p >= 0
          pi += ip;
          pf += (u64)(__fsub_rn(p, (float)ip) * 1099511627776.0f);   // exact: fraction * 2^40
        }
      }
      if (pf >> 62) { pi += pf >> 40; pf &= (1ull << 40) - 1; }
    } else {
      // a chromosome boundary inside the round (rare): per-interval atomics
      flush();
      for (int j = 0; j < 4; j++) {
        const u64 i = base + j * 256 + threadIdx.x;
        if (i < hi) {
          const int c = chrom_of_index(r.chrom_start, nchrom, i);
          const u32 e = r.end[i];
          const u32 st = (i == r.chrom_start[c]) ? 0u : r.end[i - 1];
          const float v = r.val[i];
          const float p = v < 0.0f ? 0.0f : __fmul_rn(__uint2float_rn(e - st), v);
          const u64 ip = (u64)p;
          const u64 fp = (u64)(__fsub_rn(p, (float)ip) * 1099511627776.0f);
          if (ip) atomicAdd(acc_int + c, ip);
          if (fp) atomicAdd(acc_frac + c, fp);
        }
      }
      if (base + 1024 < hi) {                          // the chromosome the next round starts in
        cur = chrom_of_index(r.chrom_start, nchrom, base + 1024);
        cs = r.chrom_start[cur]; nb = r.chrom_start[cur + 1];
      }
    }
  }
  flush();
}

void launch_rle_moment(cudaStream_t s, const DevRle& r, u64 n_upper, int nchrom,
                       u64* acc_int, u64* acc_frac) {
  if (!n_upper) return;
  u64 blocks = (n_upper + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
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
