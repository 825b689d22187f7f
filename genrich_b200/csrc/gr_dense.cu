// gr_dense.cu -- the per-base kernels: K1 delta scatter, K2 single-pass dense
// prefix sum with break compaction, K2b exact weighted-length reduction.
//
// Replaces saveInterval's diff-array writes (Genrich.c:2575-2583) and the
// sequential O(genome) loops of savePileupExpt (2239-2273) / calcFactor
// (2013-2038).  HBM-bound int32 work: no tensor cores.
#include "gr_common.cuh"
#include "gr_internal.h"
#include <stdlib.h>
#include <stdio.h>

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

// ----------------------------------------------------------------------------
// Locality pass in front of K1.  Read names arrive in queryname order, i.e. the
// records hit the 12 GB delta array at random: every RED.ADD costs a DRAM read of a
// line it shares with nobody and a write-back (ncu r01_f: 8.5 GB read + 3.2 GB
// written for 100 M atomics, 30 % of DRAM peak).  Records are therefore first moved
// into ~3000 buckets of 2^20 cells by the cell of their start (a fragment ends a
// few hundred cells further, i.e. in the same bucket); the scatter then sweeps the
// array bucket by bucket with the whole grid inside an L2-sized window, so each
// touched line is fetched and written back once, in address order.  The order of
// records inside a bucket is arbitrary -- integer atomics make the result the same.
#define BIN_CHUNK 8192            // records per CTA in the move pass

__device__ __forceinline__ u32 bin_of(const int4 r, const DevLayout& L, int shift) {
  const int c = r.x;
  if (c < 0 || c >= L.nchrom) return 0;
  const u64 off = L.off[c];
  if (off == ~0ull) return 0;
  i64 s = r.y;
  if (s < 0) s = 0;
  return (u32)((off + (u64)s) >> shift);
}

__global__ void __launch_bounds__(256)
k_bin_count(const int4* __restrict__ recs, u64 n, DevLayout L, int shift, u32 nb, u32* __restrict__ cnt) {
  extern __shared__ u32 sh[];
  for (u32 i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u32 bkt = bin_of(ld_stream_v4(recs + i), L, shift);
    if (bkt >= nb) bkt = nb - 1;
    atomicAdd(&sh[bkt], 1u);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < nb; i += blockDim.x)
    if (sh[i]) atomicAdd(cnt + i, sh[i]);
}

// exclusive scan of the bucket counts into the running cursors (one block)
__global__ void __launch_bounds__(1024)
k_bin_scan(const u32* __restrict__ cnt, u32 nb, u64* __restrict__ cursor) {
  __shared__ u64 sm_w[32];
  __shared__ u64 sm_carry;
  if (threadIdx.x == 0) sm_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u32 base = 0; base < nb; base += 1024) {
    const u32 i = base + threadIdx.x;
    const u64 v = i < nb ? cnt[i] : 0;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(GR_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm_w[w] = inc;
    __syncthreads();
    u64 wx = 0, tot = 0;
    for (int k = 0; k < 32; k++) { const u64 a = sm_w[k]; if (k < w) wx += a; tot += a; }
    const u64 carry = sm_carry;
    if (i < nb) cursor[i] = carry + wx + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) sm_carry = carry + tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_bin_move(const int4* __restrict__ recs, u64 n, DevLayout L, int shift, u32 nb,
           u64* __restrict__ cursor, int4* __restrict__ out) {
  extern __shared__ u32 sh[];                 // [nb] counts, then fill positions; [nb] reserved bases (low 32 bits) ...
  u32* cnt = sh;
  u64* base = reinterpret_cast<u64*>(sh + ((nb + 1) & ~1u));
  for (u32 i = threadIdx.x; i < nb; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  const u64 lo = (u64)blockIdx.x * BIN_CHUNK;
  const u64 hi = min(lo + BIN_CHUNK, n);
  for (u64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    u32 bkt = bin_of(recs[i], L, shift);
    if (bkt >= nb) bkt = nb - 1;
    atomicAdd(&cnt[bkt], 1u);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < nb; i += blockDim.x) {
    const u32 c = cnt[i];
    if (c) base[i] = atomicAdd(cursor + i, (u64)c);     // this CTA's slice of bucket i
    cnt[i] = 0;
  }
  __syncthreads();
  for (u64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int4 r = recs[i];                             // second read of the chunk: L1/L2
    u32 bkt = bin_of(r, L, shift);
    if (bkt >= nb) bkt = nb - 1;
    out[base[bkt] + atomicAdd(&cnt[bkt], 1u)] = r;
  }
}

void launch_scatter_binned(cudaStream_t s, const DevLayout& L, const int32_t* recs, u64 n,
                           int32_t* delta, int* err, u64* clamped, int32_t* scratch_recs,
                           u32* bin_cnt, u64* bin_cursor) {
  if (!n) return;
  int shift = 20;
  while ((L.T >> shift) + 1 > 6144) shift++;
  const u32 nb = (u32)(L.T >> shift) + 1;
  cudaMemsetAsync(bin_cnt, 0, nb * sizeof(u32), s);
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_bin_count<<<(unsigned)blocks, 256, nb * sizeof(u32), s>>>((const int4*)recs, n, L, shift, nb, bin_cnt); GR_NOTE_LAUNCH();
  k_bin_scan<<<1, 1024, 0, s>>>(bin_cnt, nb, bin_cursor); GR_NOTE_LAUNCH();
  const size_t smem = (((size_t)nb + 1) & ~(size_t)1) * sizeof(u32) + (size_t)nb * sizeof(u64);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_bin_move, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  k_bin_move<<<(unsigned)((n + BIN_CHUNK - 1) / BIN_CHUNK), 256, smem, s>>>((const int4*)recs, n, L, shift, nb, bin_cursor,
                                                                            (int4*)scratch_recs); GR_NOTE_LAUNCH();
  launch_scatter(s, L, scratch_recs, n, delta, err, clamped);
}

// ============================================================================
// K2: one pass over the dense int32 delta cells -- persistent, software-pipelined,
// warp-specialised.
//
//   grid = co-resident CTAs only (2 per SM, cooperative launch); tiles of 8192
//   cells are dealt round-robin: in round k CTA b works on tile kG+b.
//   Per CTA: 16 compute warps + 1 exchange warp.
//   * loads: cp.async (LDGSTS, 16 B/thread/op, coalesced) into a 2-stage XOR-swizzled
//     shared-memory ring; the stage of tile k is refilled with tile k+2 as soon as
//     its cells are in registers, so 1-2 tiles (32-64 KB) per CTA are always in
//     flight to HBM.
//   * A(k)  compute warps: 16 consecutive cells per thread (conflict-free LDS.128),
//           thread/warp/block scan of (sum, #breaks), break bitmap written, tile
//           aggregate handed to the exchange warp; the tile's breaks (~6 % of the
//           cells) are parked as (position, tile-local height) in a small side buffer.
//   * X(k)  exchange warp, concurrently with A(k+1), A(k+2): publishes the aggregate
//           and gathers the tile's exclusive prefix with two L2 round trips:
//             agg[tile]      the tile's own (sum, #breaks)
//             grp[k*NG + g]  total of the 32 tiles of CTA-group g in round k
//           exclusive(kG+b) = (totals of all rounds < k, kept in registers)
//                           + sum_{g' < g} grp[k,g'] + sum_{b' in group, b' < b} agg.
//           Every dependency is on AGGREGATES of the same or the previous round;
//           nothing waits for another tile's prefix, so there is no serial chain.
//   * B(k)  compute warps, two rounds later: all 512 threads convert the parked
//           breaks densely -- add the prefix, rebuild the reference float, store
//           (end, value) at consecutive global ranks (fully coalesced).
//   History (profiles/README.md): 32-wide decoupled look-back 0.92 TB/s; ticketed
//   persistent tiles 0.16 TB/s; 320-wide look-back 0.77 TB/s; two-level exchange,
//   prefix awaited in place 1.03 TB/s; + exchange warp, emission one round late
//   1.84 TB/s; + dense conversion of the breaks 2.1 TB/s (33 % of warp samples still
//   parked behind the exchange: CTAs drift by more than one round).
//
// A break closes an interval at chromosome position j iff 1 <= j < len and
// delta[j] != 0, or j == len (Genrich.c:2241, 2268); its value is the running sum
// BEFORE delta[j] is added (2245), rebuilt as the reference float.
// Running sums are kept modulo 2^32: every true prefix fits in int32.
#define SC_ITEMS 16
#define SC_NSTAGE 2
#define BAR_COMPUTE 1
#define BAR_AGG 2                 // + slot
#define BAR_PREFIX 8              // + slot

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory");
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

// status words, all {flag|sum32, flag|count}, zeroed before the launch
struct ScanStatus { ulonglong2* agg; ulonglong2* grp; u32 ngroups; };

struct TileMeta { u64 off; u32 len; int c; bool act; };
__device__ __forceinline__ TileMeta tile_meta(const DevLayout& L, u32 tile, u32 ntiles, u32 tile_cells) {
  TileMeta m;
  m.c = 0; m.off = 0; m.len = 0; m.act = false;
  if (tile < ntiles) {
    m.c = L.blk2chrom[((u64)tile * tile_cells) >> GR_BLOCK_SHIFT];
    m.off = L.off[m.c];
    m.len = L.len[m.c];
    m.act = (L.flags[m.c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
  }
  return m;
}

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

// SC_LAG: rounds between A(k) and B(k); SC_NSLOT = SC_LAG + 1 side buffers / hand-over slots;
// SC_SIDE_CAP: parked breaks per tile (8 B each) -- denser tiles take the dense path.
// SC_CT: compute threads per CTA (16 cells each -> SC_CT*16 cells per tile), + one exchange warp.
template <int SC_CT, int SC_LAG, int SC_SIDE_CAP>
__global__ void __launch_bounds__(SC_CT + 32, SC_CT == 512 ? 2 : 4)
k_dense_scan(int32_t* __restrict__ delta, DevLayout L, ScanStatus S, DevRle out,
             u32* __restrict__ bitmap, int* __restrict__ err, u32 ntiles, int zero_after) {
  constexpr int SC_NSLOT = SC_LAG + 1;
  constexpr int SC_WARPS = SC_CT / 32, SC_THREADS = SC_CT + 32, SC_STAGE_INT4 = SC_CT * 4;
  constexpr u32 TILE = SC_CT * 16;                     // cells per tile (a divisor of GR_BLOCK_SLOTS)
  extern __shared__ int4 sm_x[];                       // SC_NSTAGE stages, then SC_NSLOT side buffers
  __shared__ u32 sm_wsum[SC_WARPS], sm_wcnt[SC_WARPS];
  __shared__ u32 sm_agg_sum[SC_NSLOT], sm_agg_cnt[SC_NSLOT];
  __shared__ u32 sm_ex_sum[SC_NSLOT];
  __shared__ u64 sm_ex_cnt[SC_NSLOT];
  __shared__ u32 sm_q[SC_NSLOT][5];                    // jb, tile, chrom, first|live<<1, #breaks
  __shared__ u32 sm_bm[SC_CT / 2];                     // the tile's break bitmap (16 words per warp)
  __shared__ unsigned char sm_list[SC_WARPS * 128];    // per warp: its non-zero chunks
  __shared__ float4 sm_lut[120];                       // height mod 120 -> fractional float terms

  const int tid = threadIdx.x, lane = tid & 31;
  const u32 G = gridDim.x, b = blockIdx.x;
  int2* side = reinterpret_cast<int2*>(sm_x + SC_NSTAGE * SC_STAGE_INT4);

  // ------------------------------------------------------------ exchange warp
  if (tid >= SC_CT) {
    const u32 g = b >> 5, j = b & 31, ng = S.ngroups;
    u32 rnd_s = 0;                                     // totals of all rounds before k
    u64 rnd_c = 0;
    u32 k = 0;
    for (u32 tile = b; tile < ntiles; tile += G, k++) {
      const int slot = k % SC_NSLOT;
      named_sync(BAR_AGG + slot, 64);                  // aggregate of tile k is in shared memory
      const u32 agg_s = sm_agg_sum[slot], agg_c = sm_agg_cnt[slot];
      if (lane == 0) st_status(S.agg + tile, 1, agg_s, agg_c);
      const u32 last_b = min(G - 1, ntiles - 1 - k * G);
      const bool need_a = (u32)lane < j, need_g = (u32)lane < g, need_p = k > 0 && (u32)lane < ng;
      const ulonglong2* pa = S.agg + (tile - j) + lane;
      const ulonglong2* pg = S.grp + (u64)k * ng + lane;
      const ulonglong2* pp = S.grp + (u64)(k - 1) * ng + lane;   // only dereferenced when k > 0
      // One polling loop, all loads of a round issued together (one L2 round trip when
      // everything is there).  The group's last tile publishes the group total as soon
      // as the group's own aggregates are in -- it must NOT wait for the totals of
      // earlier groups, or the groups of a round serialise (measured: 35 polls per tile).
      const bool is_last = b == min(32u * g + 31u, last_b);
      ulonglong2 va, vg, vp;
      va.x = va.y = vg.x = vg.y = vp.x = vp.y = 0;
      bool ok_a = !need_a, ok_o = !(need_g || need_p), published = !is_last;
      u32 s_in = 0;
      u64 c_in = 0;
      for (;;) {
        if (!ok_a) {
          va = ld_status(pa);
          ok_a = (va.x >> 62) == 1 && (va.y >> 62) == 1;
        }
        if (!ok_o) {
          bool okg = true, okp = true;
          if (need_g) { vg = ld_status(pg); okg = (vg.x >> 62) == 1 && (vg.y >> 62) == 1; }
          if (need_p) { vp = ld_status(pp); okp = (vp.x >> 62) == 1 && (vp.y >> 62) == 1; }
          ok_o = okg && okp;
        }
        const bool all_a = __all_sync(GR_FULL, ok_a);
        if (all_a && !published) {
          s_in = need_a ? (u32)va.x : 0u;
          c_in = need_a ? (va.y & GR_LB_PAYLOAD) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s_in += __shfl_xor_sync(GR_FULL, s_in, o);
            c_in += __shfl_xor_sync(GR_FULL, c_in, o);
          }
          if (lane == 0) st_status(S.grp + (u64)k * ng + g, 1, s_in + agg_s, c_in + agg_c);
          published = true;
        }
        if (all_a && __all_sync(GR_FULL, ok_o)) break;
      }
      if (!is_last) {
        s_in = need_a ? (u32)va.x : 0u;
        c_in = need_a ? (va.y & GR_LB_PAYLOAD) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s_in += __shfl_xor_sync(GR_FULL, s_in, o);
          c_in += __shfl_xor_sync(GR_FULL, c_in, o);
        }
      }
      u32 s_g = need_g ? (u32)vg.x : 0u, s_p = need_p ? (u32)vp.x : 0u;
      u64 c_g = need_g ? (vg.y & GR_LB_PAYLOAD) : 0ull, c_p = need_p ? (vp.y & GR_LB_PAYLOAD) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_g += __shfl_xor_sync(GR_FULL, s_g, o);
        c_g += __shfl_xor_sync(GR_FULL, c_g, o);
        s_p += __shfl_xor_sync(GR_FULL, s_p, o);
        c_p += __shfl_xor_sync(GR_FULL, c_p, o);
      }
      rnd_s += s_p; rnd_c += c_p;
      if (lane == 0) {
        sm_ex_sum[slot] = rnd_s + s_g + s_in;
        sm_ex_cnt[slot] = rnd_c + c_g + c_in;
      }
      __syncwarp();
      named_arrive(BAR_PREFIX + slot, SC_THREADS);     // prefix of tile k is in shared memory
    }
    return;
  }

  // ------------------------------------------------------------ compute warps
  const int w = tid >> 5;
  const int4* src = reinterpret_cast<const int4*>(delta) + (u64)b * (TILE / 4) + tid;   // read side
  const u64 src_step = (u64)G * (TILE / 4);
  auto issue = [&](bool on, int stage, const int4* from) {   // chunk q*512+tid (16 B), same place in the stage
    if (on) {
      int4* dst = sm_x + stage * SC_STAGE_INT4 + tid;
#pragma unroll
      for (int q = 0; q < 4; q++) cp_async16(dst + q * SC_CT, from + q * SC_CT);
    }
    cp_async_commit();
  };
  issue(b < ntiles, 0, src);
  issue(b + G < ntiles, 1, src + src_step);
  TileMeta meta = tile_meta(L, b, ntiles, TILE);
  if (tid < SC_NSLOT) sm_q[tid][3] = 0;
  units_lut_fill(sm_lut, tid, SC_CT);                  // visible after the first BAR_COMPUTE sync

  // B: convert and store the parked breaks of the tile in `slot` (all 512 threads, dense)
  auto finish = [&](int slot) {
    const u32 jb_ = sm_q[slot][0], tile_ = sm_q[slot][1], fl = sm_q[slot][3];
    const int c_ = (int)sm_q[slot][2];
    const u32 cnt_ = sm_q[slot][4];
    named_sync(BAR_PREFIX + slot, SC_THREADS);
    const u32 ex_sum = sm_ex_sum[slot];
    const u64 ex_cnt = sm_ex_cnt[slot];
    if (tid == 0) {
      if (fl & 1) {
        out.chrom_start[c_] = ex_cnt;
        if (ex_sum != 0) atomicOr(err, GR_DE_TAIL);    // previous chromosome did not return to 0 (2283-2289)
      }
      if (tile_ == ntiles - 1) {
        *out.total = ex_cnt + cnt_;
        out.chrom_start[L.nchrom] = ex_cnt + cnt_;
      }
    }
    bool neg = false;
    for (u32 n = tid; n < cnt_; n += SC_CT) {          // thread n <-> n-th break of the tile: coalesced stores
      const int2 e = side[slot * SC_SIDE_CAP + n];
      const int N = (int)(ex_sum + (u32)e.y);
      neg |= N < 0;
      out.end[ex_cnt + n] = jb_ + (u32)e.x;
      out.val[ex_cnt + n] = units_to_val_lut(sm_lut, N < 0 ? 0 : N);
      // every break of an interior tile is a non-zero cell and vice versa: clearing them
      // leaves the whole delta array zero for the next sample (no 4 B/bp memset)
      if (zero_after) delta[(u64)tile_ * TILE + (u32)e.x] = 0;
    }
    if (neg) atomicOr(err, GR_DE_PILE);                // ERRPILE 1921, 1969
  };

  u32 k = 0;
  for (u32 tile = b; tile < ntiles; tile += G, k++) {
    const int slot = k % SC_NSLOT;
    // chromosome of the tile after this one: first hop now, second hop after the scan
    const int c_next = tile + G < ntiles ? L.blk2chrom[((u64)(tile + G) * TILE) >> GR_BLOCK_SHIFT] : 0;
    cp_async_wait<1>();
    named_sync(BAR_COMPUTE, SC_CT);                    // tile k is in shared memory (all threads' copies)
    const int4* stage = sm_x + (k & 1) * SC_STAGE_INT4;
    const u64 tbase = (u64)tile * TILE;
    const u32 jb = (u32)(tbase - meta.off);            // chromosome position of the tile's first cell
    const u32 len = meta.len;
    const bool interior = jb >= 1 && (u64)jb + TILE <= (u64)len;
    bool fast = interior;

    // dense scan of the thread's 16 cells (chromosome ends, over-full tiles)
    int d[SC_ITEMS];
    u32 run = 0, m = 0, cnt = 0, wi_sum = 0, wi_cnt = 0;
    auto dense_part1 = [&]() {
      sc_load_items(stage, tid, d);
      run = 0; m = 0;
      const u32 j0 = jb + tid * SC_ITEMS;
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        run += (u32)d[i];
        const u32 jj = j0 + i;
        const bool brk = (jj == len) || (d[i] != 0 && jj >= 1 && jj < len);
        m |= (brk ? 1u : 0u) << i;
      }
      if (!meta.act) m = 0;
      cnt = __popc(m);
      wi_sum = warp_incl_scan_u32(run, lane);
      wi_cnt = warp_incl_scan_u32(cnt, lane);
      if (lane == 31) { sm_wsum[w] = wi_sum; sm_wcnt[w] = wi_cnt; }
      const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
      if (!(lane & 1)) bitmap[(tbase >> 5) + (tid >> 1)] = m | (hi << 16);
    };

    const int4* wst = stage + w * 128;                 // this warp's 512 cells = 128 chunks of 16 B
    unsigned char* wlist = sm_list + w * 128;          // indices of its non-zero chunks, in order
    u32 nnz = 0;
    if (fast) {
      // ---- A(k), sparse-aware: ~94 % of the cells are zero; they neither move the running
      // sum nor break an interval.  Each warp tests its 128 chunks with four coalesced
      // LDS.128 + ballots, lists the non-zero ones, and from here on only they cost work.
      u32 cprev = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int4 x = wst[r * 32 + lane];
        const bool nz = (x.x | x.y | x.z | x.w) != 0;
        const u32 M = __ballot_sync(GR_FULL, nz);
        if (nz) wlist[cprev + __popc(M & ((1u << lane) - 1))] = (unsigned char)(r * 32 + lane);
        cprev += __popc(M);
      }
      nnz = cprev;
      if (lane < 16) sm_bm[w * 16 + lane] = 0;
      __syncwarp();
      // phase 1: the warp's totals (sum of deltas, number of non-zero cells)
      u32 ts = 0, tc = 0;
      for (u32 base = 0; base < nnz; base += 32) {
        const u32 n = base + lane;
        int4 x = make_int4(0, 0, 0, 0);
        if (n < nnz) x = wst[wlist[n]];
        const u32 c4 = (x.x != 0) + (x.y != 0) + (x.z != 0) + (x.w != 0);
        ts += __reduce_add_sync(GR_FULL, (u32)x.x + (u32)x.y + (u32)x.z + (u32)x.w);
        tc += __reduce_add_sync(GR_FULL, c4);
      }
      if (lane == 0) { sm_wsum[w] = ts; sm_wcnt[w] = tc; }
    } else
      dense_part1();
    named_sync(BAR_COMPUTE, SC_CT);                    // warp totals are in shared memory
    // exclusive prefix over the 16 warp totals: lanes 0..15 scan them, everyone picks its warp's
    u32 vs = lane < SC_WARPS ? sm_wsum[lane] : 0u, vc = lane < SC_WARPS ? sm_wcnt[lane] : 0u;
    u32 is_ = warp_incl_scan_u32(vs, lane), ic_ = warp_incl_scan_u32(vc, lane);
    u32 t_sum = __shfl_sync(GR_FULL, is_, SC_WARPS - 1), t_cnt = __shfl_sync(GR_FULL, ic_, SC_WARPS - 1);
    if (fast && t_cnt > SC_SIDE_CAP) {                 // > 18.75 % of the tile's cells are breaks: dense path
      named_sync(BAR_COMPUTE, SC_CT);                  // everyone has read the totals
      dense_part1();
      named_sync(BAR_COMPUTE, SC_CT);
      vs = lane < SC_WARPS ? sm_wsum[lane] : 0u; vc = lane < SC_WARPS ? sm_wcnt[lane] : 0u;
      is_ = warp_incl_scan_u32(vs, lane); ic_ = warp_incl_scan_u32(vc, lane);
      t_sum = __shfl_sync(GR_FULL, is_, SC_WARPS - 1); t_cnt = __shfl_sync(GR_FULL, ic_, SC_WARPS - 1);
      fast = false;
    }
    const bool first = tbase == meta.off;
    if (w == 0) {
      if (lane == 0) {
        sm_agg_sum[slot] = t_sum; sm_agg_cnt[slot] = t_cnt;
        if (fast) {
          sm_q[slot][0] = jb; sm_q[slot][1] = tile; sm_q[slot][2] = (u32)meta.c; sm_q[slot][3] = (first ? 1u : 0u) | 2u;
          sm_q[slot][4] = t_cnt;
        } else
          sm_q[slot][3] = 0;
      }
      __syncwarp();
      named_arrive(BAR_AGG + slot, 64);                // hand over to the exchange warp
    }
    if (fast) {
      // phase 2: park the breaks at their rank inside the tile: (position, tile-local height)
      const u32 wx_sum = __shfl_sync(GR_FULL, is_ - vs, w), wx_cnt = __shfl_sync(GR_FULL, ic_ - vc, w);
      int2* sb = side + slot * SC_SIDE_CAP;
      u32 carry_s = wx_sum, carry_c = wx_cnt;
      for (u32 base = 0; base < nnz; base += 32) {
        const u32 n = base + lane;
        const bool on = n < nnz;
        int4 x = make_int4(0, 0, 0, 0);
        u32 q = 0;
        if (on) { q = wlist[n]; x = wst[q]; }
        const u32 m4 = (x.x != 0 ? 1u : 0u) | (x.y != 0 ? 2u : 0u) | (x.z != 0 ? 4u : 0u) | (x.w != 0 ? 8u : 0u);
        const u32 c4 = __popc(m4);
        const u32 s1 = (u32)x.x, s2 = s1 + (u32)x.y, s3 = s2 + (u32)x.z, s4 = s3 + (u32)x.w;
        const u32 inc_s = warp_incl_scan_u32(s4, lane), inc_c = warp_incl_scan_u32(c4, lane);
        if (on) {
          const u32 ex_s = carry_s + inc_s - s4;       // tile-local running sum before this chunk
          int2* e = sb + (carry_c + inc_c - c4);
          const int p0 = w * 512 + (int)(q * 4);
          if (m4 & 1u) *e++ = make_int2(p0, (int)ex_s);
          if (m4 & 2u) *e++ = make_int2(p0 + 1, (int)(ex_s + s1));
          if (m4 & 4u) *e++ = make_int2(p0 + 2, (int)(ex_s + s2));
          if (m4 & 8u) *e++ = make_int2(p0 + 3, (int)(ex_s + s3));
          atomicOr(&sm_bm[w * 16 + (q >> 3)], m4 << ((q & 7) * 4));
        }
        carry_s += __shfl_sync(GR_FULL, inc_s, 31);
        carry_c += __shfl_sync(GR_FULL, inc_c, 31);
      }
      __syncwarp();
      if (lane < 16) bitmap[(tbase >> 5) + w * 16 + lane] = sm_bm[w * 16 + lane];
      named_sync(BAR_COMPUTE, SC_CT);                  // the stage has been consumed: refill it
      src += src_step;
      issue(tile + 2 * G < ntiles, k & 1, src + src_step);
    }
    if (!fast) {
      // dense path: wait for this tile's prefix and write its breaks from registers
      const u32 wx_sum = __shfl_sync(GR_FULL, is_ - vs, w), wx_cnt = __shfl_sync(GR_FULL, ic_ - vc, w);
      named_sync(BAR_PREFIX + slot, SC_THREADS);
      const u32 ex_sum = sm_ex_sum[slot];
      const u64 ex_cnt = sm_ex_cnt[slot];
      if (tid == 0) {
        if (first) {
          out.chrom_start[meta.c] = ex_cnt;
          if (ex_sum != 0) atomicOr(err, GR_DE_TAIL);
        }
        if (tile == ntiles - 1) {
          *out.total = ex_cnt + t_cnt;
          out.chrom_start[L.nchrom] = ex_cnt + t_cnt;
        }
      }
      if (m) {
        u32 rr = ex_sum + wx_sum + (wi_sum - run);
        u64 rank = ex_cnt + wx_cnt + (wi_cnt - cnt);
        const u32 j0 = jb + tid * SC_ITEMS;
        bool neg = false;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
          if (m & (1u << i)) {
            const int N = (int)rr;
            neg |= N < 0;
            out.end[rank] = j0 + i;
            out.val[rank] = units_to_val_lut(sm_lut, N < 0 ? 0 : N);
            rank++;
          }
          rr += (u32)d[i];
        }
        if (neg) atomicOr(err, GR_DE_PILE);
      }
      if (zero_after) {
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++)
          if (d[i] != 0) delta[tbase + tid * SC_ITEMS + i] = 0;
      }
      named_sync(BAR_COMPUTE, SC_CT);                  // everyone is done with the stage
      src += src_step;
      issue(tile + 2 * G < ntiles, k & 1, src + src_step);
    }
    TileMeta meta_next;
    meta_next.c = c_next;
    meta_next.off = L.off[c_next];
    meta_next.len = L.len[c_next];
    meta_next.act = (L.flags[c_next] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);

    // ---- B(k-LAG): the prefix of that tile has had two rounds to arrive
    if (k >= SC_LAG) {
      const int ps = (k - SC_LAG) % SC_NSLOT;
      if (sm_q[ps][3] & 2u) finish(ps);
    }
    meta = meta_next;
  }
  named_sync(BAR_COMPUTE, SC_CT);
  for (u32 kk = (k >= SC_LAG ? k - SC_LAG : 0); kk < k; kk++) {
    const int ps = kk % SC_NSLOT;
    if (sm_q[ps][3] & 2u) finish(ps);
  }
  cp_async_wait<0>();
}


// ----------------------------------------------------------------------------
// K2, warp-autonomous form.  Same tiling, same exchange, same results as above, but
// the 16 compute warps of a CTA never meet at a CTA barrier:
//   * every warp streams ITS OWN 512 cells of each tile through a private 2-stage
//     cp.async ring (4 x 512 contiguous bytes per tile, waited on with __syncwarp only);
//   * one pass per tile: non-zero 16-byte chunks are listed (ballot), counted, the warp
//     reserves room in the tile's side pool (one shared-memory atomic) and parks its
//     breaks as (position, WARP-local height); its totals go to shared memory and it
//     bar.arrive's -- it does not wait for the block prefix;
//   * the exchange warp sums the 16 warp totals, publishes / gathers as before, and
//     hands each warp its own exclusive prefix (sum, rank); completion is signalled
//     through an mbarrier per slot, so a warp that converts its parked breaks LAG
//     tiles later waits alone, and only if the prefix is not there yet.
// Slot ring: a pool slot is rewritten at tile k+NSLOT by warps that have converted
// tile k+NSLOT-LAG, which the exchange releases only after EVERY warp arrived for that
// tile, i.e. after every warp converted tile k+NSLOT-2*LAG: NSLOT = 2*LAG is safe.
#define SCW_NONE 0xffffffffu
#ifdef GR_SCAN_PROF
__device__ unsigned long long g_scan_prof[8];
__device__ __forceinline__ u64 gtime() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PROF_T(v) const u64 v = gtime()
#define PROF_ADD(i, x) do { if (lane == 0) atomicAdd(&g_scan_prof[i], (unsigned long long)(x)); } while (0)
#else
#define PROF_T(v)
#define PROF_ADD(i, x)
#endif
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n"
      "D_%=:\n\t}" :: "r"(a), "r"(parity) : "memory");
}

template <int LAG, int CAP, int NX>
__global__ void __launch_bounds__(512 + 32 * NX, 2)
k_dense_scan_w(int32_t* __restrict__ delta, DevLayout L, ScanStatus S, DevRle out,
               u32* __restrict__ bitmap, int* __restrict__ err, u32 ntiles, int zero_after) {
  constexpr int NSLOT = 2 * LAG, CT = 512, NW = 16;
  constexpr u32 TILE = 8192;
  extern __shared__ int4 sm_x[];                       // per warp 2 stages x 128 chunks, then NSLOT pools
  __shared__ u32 sm_wsum[NSLOT][NW], sm_wcnt[NSLOT][NW], sm_woff[NSLOT][NW];
  __shared__ u32 sm_pre_sum[NSLOT][NW];
  __shared__ u64 sm_pre_cnt[NSLOT][NW];
  __shared__ u32 sm_alloc[NSLOT];
  __shared__ u32 sm_rs[4], sm_rready;                  // running totals of completed rounds (exchange warps)
  __shared__ u64 sm_rc[4];
  __shared__ u64 sm_bar[NSLOT];
  __shared__ u32 sm_bm[NW * 16];
  __shared__ unsigned char sm_list[NW * 128];
  __shared__ float4 sm_lut[120];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u32 G = gridDim.x, b = blockIdx.x;
  int2* pool = reinterpret_cast<int2*>(sm_x + NW * 2 * 128);

  units_lut_fill(sm_lut, tid, 512 + 32 * NX);
  if (tid < NSLOT) { sm_alloc[tid] = 0; mbar_init(&sm_bar[tid], 1); }
  if (tid == 0) { sm_rs[0] = 0; sm_rc[0] = 0; sm_rready = 0; }
  __syncthreads();

  // ------------------------------------------------------------ exchange warp
  // NX exchange warps take the rounds in turn (round k -> warp k % NX), so NX exchanges are
  // in flight per CTA.  One hop: a tile's aggregate is ONE 64-bit status word
  // (flag:2 | #breaks:30 | sum:32); the exchange of tile (k, b) reads the words of ALL
  // tiles of round k -- those before b give the exclusive prefix inside the round, all of
  // them the round total -- and the total of the earlier rounds is handed from round to
  // round inside the CTA through shared memory.  (The two-level agg -> group-total scheme
  // needed two dependent store->poll hops, ~8 us under load: three tiles of slack did not
  // cover it and half of all warp samples sat waiting for a prefix.)
  if (w >= NW) {
    u64* agg64 = reinterpret_cast<u64*>(S.agg);
    for (u32 k = (u32)(w - NW); (u64)b + (u64)k * G < ntiles; k += NX) {
      const u32 tile = b + k * G;
      const int slot = k % NSLOT;
      // what lane 0 needs for the chromosome bookkeeping, fetched ahead of the barrier
      const int c_t = L.blk2chrom[tile];
      const u64 off_t = L.off[c_t];
      const u32 nb = min(G, ntiles - k * G);           // tiles in this round
      const u64* base = agg64 + (u64)k * G;
      PROF_T(t0);
      named_sync(BAR_AGG + slot, 544);                 // (16 compute warps + this one) all warp totals are in
      PROF_T(t1);
      const u32 ws = lane < NW ? sm_wsum[slot][lane] : 0u, wc = lane < NW ? sm_wcnt[slot][lane] : 0u;
      if (lane == 0) sm_alloc[slot] = 0;
      const u32 is_ = warp_incl_scan_u32(ws, lane), ic_ = warp_incl_scan_u32(wc, lane);
      const u32 agg_s = __shfl_sync(GR_FULL, is_, NW - 1), agg_c = __shfl_sync(GR_FULL, ic_, NW - 1);
      if (lane == 0) st_relaxed_u64(agg64 + tile, (1ull << 62) | ((u64)agg_c << 32) | agg_s);
      u32 pre_s = 0, pre_c = 0, tot_s = 0, tot_c = 0;
      for (u32 c0 = 0; c0 < nb; c0 += 256) {           // 8 status words per lane and pass
        u64 v[8];
        u32 pending = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (c0 + i * 32 + lane < nb) pending |= 1u << i;
        for (;;) {
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (pending & (1u << i)) v[i] = ld_relaxed_u64(base + c0 + i * 32 + lane);
#pragma unroll
          for (int i = 0; i < 8; i++)
            if ((pending & (1u << i)) && (v[i] >> 62) == 1) {
              pending &= ~(1u << i);
              const u32 vs = (u32)v[i], vc = (u32)(v[i] >> 32) & 0x3fffffffu;
              tot_s += vs; tot_c += vc;
              if (c0 + i * 32 + lane < b) { pre_s += vs; pre_c += vc; }
            }
          PROF_ADD(3, 1);
          if (__all_sync(GR_FULL, pending == 0)) break;
        }
      }
      PROF_T(t2);
      pre_s = __reduce_add_sync(GR_FULL, pre_s); pre_c = __reduce_add_sync(GR_FULL, pre_c);
      tot_s = __reduce_add_sync(GR_FULL, tot_s); tot_c = __reduce_add_sync(GR_FULL, tot_c);
      // totals of the rounds before k: from the warp that exchanged round k-1
      while (*(volatile u32*)&sm_rready < k) { }
      __threadfence_block();
      PROF_T(t3);
      PROF_ADD(0, t1 - t0); PROF_ADD(1, t2 - t1); PROF_ADD(2, t3 - t2); PROF_ADD(4, 1);
      const u32 r_s = sm_rs[k & 3];
      const u64 r_c = sm_rc[k & 3];
      if (lane == 0) {
        sm_rs[(k + 1) & 3] = r_s + tot_s;
        sm_rc[(k + 1) & 3] = r_c + tot_c;
        __threadfence_block();
        *(volatile u32*)&sm_rready = k + 1;
      }
      const u32 ex_s = r_s + pre_s;
      const u64 ex_c = r_c + pre_c;
      if (lane < NW) {
        sm_pre_sum[slot][lane] = ex_s + (is_ - ws);
        sm_pre_cnt[slot][lane] = ex_c + (u64)(ic_ - wc);
      }
      if (lane == 0) {
        if ((u64)tile * TILE == off_t) {               // first tile of a chromosome
          out.chrom_start[c_t] = ex_c;
          if (ex_s != 0) atomicOr(err, GR_DE_TAIL);    // previous chromosome did not return to 0 (2283-2289)
        }
        if (tile == ntiles - 1) {
          *out.total = ex_c + agg_c;
          out.chrom_start[L.nchrom] = ex_c + agg_c;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm_bar[slot]);       // release: the 16 prefixes of tile k are visible
    }
    return;
  }

  // ------------------------------------------------------------ compute warps
  int4* wring = sm_x + w * 256;                        // this warp's two stages of 128 chunks
  unsigned char* wlist = sm_list + w * 128;
  const int4* src = reinterpret_cast<const int4*>(delta) + (u64)b * (TILE / 4) + w * 128 + lane;
  const u64 src_step = (u64)G * (TILE / 4);
  auto issue = [&](bool on, int stage, const int4* from) {
    if (on) {
      int4* dst = wring + stage * 128 + lane;
#pragma unroll
      for (int r = 0; r < 4; r++) cp_async16(dst + r * 32, from + r * 32);
    }
    cp_async_commit();
  };
  issue(b < ntiles, 0, src);
  issue(b + G < ntiles, 1, src + src_step);
  TileMeta meta = tile_meta(L, b, ntiles, TILE);
  u32 jb_hist[LAG];                                    // chromosome position of the first cell of tiles k-1 .. k-LAG
#pragma unroll
  for (int i = 0; i < LAG; i++) jb_hist[i] = 0;
  const u32 lt_mask = (1u << lane) - 1;

  // B: convert this warp's parked breaks of tile kk (dense, coalesced within the warp's run)
  auto finish = [&](u32 kk, u32 jb_) {
    const int slot = kk % NSLOT;
    const u32 off = sm_woff[slot][w];
    if (off == SCW_NONE) return;
    const u32 tc = sm_wcnt[slot][w];
    PROF_T(tw0);
    mbar_wait(&sm_bar[slot], (kk / NSLOT) & 1);
    PROF_T(tw1);
    if (w == 0) { PROF_ADD(5, tw1 - tw0); PROF_ADD(6, 1); }
    const u32 ps = sm_pre_sum[slot][w];
    const u64 pc = sm_pre_cnt[slot][w];
    const u64 tb = (u64)(b + kk * G) * TILE;
    const int2* sb = pool + slot * CAP + off;
    bool neg = false;
    for (u32 n = lane; n < tc; n += 32) {
      const int2 e = sb[n];
      const int N = (int)(ps + (u32)e.y);
      neg |= N < 0;
      out.end[pc + n] = jb_ + (u32)e.x;
      out.val[pc + n] = units_to_val_lut(sm_lut, N < 0 ? 0 : N);
      // every break of an interior tile is a non-zero cell and vice versa: clearing them
      // leaves the whole delta array zero for the next sample (no 4 B/bp memset)
      if (zero_after) delta[tb + (u32)e.x] = 0;
    }
    if (neg) atomicOr(err, GR_DE_PILE);                // ERRPILE 1921, 1969
  };

  u32 k = 0;
  for (u32 tile = b; tile < ntiles; tile += G, k++) {
    const int slot = k % NSLOT;
    const int c_next = tile + G < ntiles ? L.blk2chrom[tile + G] : 0;
    if (k >= LAG) finish(k - LAG, jb_hist[LAG - 1]);
    cp_async_wait<1>();
    __syncwarp();                                      // the warp's 512 cells of tile k are in its stage
    const int4* wst = wring + (k & 1) * 128;
    const u64 tbase = (u64)tile * TILE;
    const u32 jb = (u32)(tbase - meta.off);
    const u32 len = meta.len;
    bool fast = jb >= 1 && (u64)jb + TILE <= (u64)len;
    u32 nnz = 0, tc = 0, off = 0;
    if (fast) {
      // list the non-zero chunks, count the non-zero cells
      u32 cc = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int4 x = wst[r * 32 + lane];
        const u32 c4 = min((u32)x.x, 1u) + min((u32)x.y, 1u) + min((u32)x.z, 1u) + min((u32)x.w, 1u);
        const u32 M = __ballot_sync(GR_FULL, c4 != 0);
        if (c4) wlist[nnz + __popc(M & lt_mask)] = (unsigned char)(r * 32 + lane);
        nnz += __popc(M);
        cc += c4;
      }
      tc = __reduce_add_sync(GR_FULL, cc);
      if (lane == 0) off = atomicAdd(&sm_alloc[slot], tc);
      off = __shfl_sync(GR_FULL, off, 0);
      if (off + tc > (u32)CAP) fast = false;           // side pool full: this warp goes the dense way
    }
    if (fast) {
      if (lane < 16) sm_bm[w * 16 + lane] = 0;
      __syncwarp();
      int2* sb = pool + slot * CAP;
      u32 carry_s = 0, carry_c = off;
      for (u32 base = 0; base < nnz; base += 32) {
        const u32 n = base + lane;
        const bool on = n < nnz;
        int4 x = make_int4(0, 0, 0, 0);
        u32 q = 0;
        if (on) { q = wlist[n]; x = wst[q]; }
        const u32 m4 = (x.x != 0 ? 1u : 0u) | (x.y != 0 ? 2u : 0u) | (x.z != 0 ? 4u : 0u) | (x.w != 0 ? 8u : 0u);
        const u32 c4 = __popc(m4);
        const u32 s1 = (u32)x.x, s2 = s1 + (u32)x.y, s3 = s2 + (u32)x.z, s4 = s3 + (u32)x.w;
        const u32 inc_s = warp_incl_scan_u32(s4, lane), inc_c = warp_incl_scan_u32(c4, lane);
        if (on) {
          const u32 ex_s = carry_s + inc_s - s4;       // warp-local running sum before this chunk
          int2* e = sb + (carry_c + inc_c - c4);
          const int p0 = w * 512 + (int)(q * 4);
          if (m4 & 1u) *e++ = make_int2(p0, (int)ex_s);
          if (m4 & 2u) *e++ = make_int2(p0 + 1, (int)(ex_s + s1));
          if (m4 & 4u) *e++ = make_int2(p0 + 2, (int)(ex_s + s2));
          if (m4 & 8u) *e++ = make_int2(p0 + 3, (int)(ex_s + s3));
          atomicOr(&sm_bm[w * 16 + (q >> 3)], m4 << ((q & 7) * 4));
        }
        carry_s += __shfl_sync(GR_FULL, inc_s, 31);
        carry_c += __shfl_sync(GR_FULL, inc_c, 31);
      }
      if (lane == 0) { sm_wsum[slot][w] = carry_s; sm_wcnt[slot][w] = tc; sm_woff[slot][w] = off; }
      __syncwarp();
      named_arrive(BAR_AGG + slot, 544);               // hand the totals to the exchange warp; do not wait
      if (lane < 16) bitmap[(tbase >> 5) + w * 16 + lane] = sm_bm[w * 16 + lane];
    } else {
      // dense: chromosome ends, inactive chromosomes, over-full tiles.  16 consecutive cells per lane.
      int d[SC_ITEMS];
      sc_load_items(wst, lane, d);
      u32 run = 0, m = 0;
      const u32 j0 = jb + w * 512 + lane * SC_ITEMS;
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        run += (u32)d[i];
        const u32 jj = j0 + i;
        const bool brk = (jj == len) || (d[i] != 0 && jj >= 1 && jj < len);
        m |= (brk ? 1u : 0u) << i;
      }
      if (!meta.act) m = 0;
      const u32 cnt = __popc(m);
      const u32 wi_sum = warp_incl_scan_u32(run, lane), wi_cnt = warp_incl_scan_u32(cnt, lane);
      if (lane == 31) { sm_wsum[slot][w] = wi_sum; sm_wcnt[slot][w] = wi_cnt; sm_woff[slot][w] = SCW_NONE; }
      __syncwarp();
      named_arrive(BAR_AGG + slot, 544);
      const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
      if (!(lane & 1)) bitmap[(tbase >> 5) + w * 16 + (lane >> 1)] = m | (hi << 16);
      mbar_wait(&sm_bar[slot], (k / NSLOT) & 1);       // this tile's prefix, synchronously (rare)
      if (m) {
        u32 rr = sm_pre_sum[slot][w] + (wi_sum - run);
        u64 rank = sm_pre_cnt[slot][w] + (wi_cnt - cnt);
        bool neg = false;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
          if (m & (1u << i)) {
            const int N = (int)rr;
            neg |= N < 0;
            out.end[rank] = j0 + i;
            out.val[rank] = units_to_val_lut(sm_lut, N < 0 ? 0 : N);
            rank++;
          }
          rr += (u32)d[i];
        }
        if (neg) atomicOr(err, GR_DE_PILE);
      }
      if (zero_after) {
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++)
          if (d[i] != 0) delta[tbase + w * 512 + lane * SC_ITEMS + i] = 0;
      }
    }
    __syncwarp();                                      // every lane is done with the stage: refill it
    src += src_step;
    issue(tile + 2 * G < ntiles, k & 1, src + src_step);
#pragma unroll
    for (int i = LAG - 1; i > 0; i--) jb_hist[i] = jb_hist[i - 1];
    jb_hist[0] = jb;
    meta.c = c_next;
    meta.off = L.off[c_next];
    meta.len = L.len[c_next];
    meta.act = (L.flags[c_next] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
  }
  // drain: tiles k-LAG .. k-1 (jb_hist[LAG-1] is the oldest)
#pragma unroll
  for (int i = LAG - 1; i >= 0; i--)
    if (k >= (u32)(i + 1)) finish(k - 1 - i, jb_hist[i]);
  cp_async_wait<0>();
}


// ============================================================================
// K2, streaming form: no cross-warp dependency at all inside the 4 B/cell pass.
//
// Why: in every single-pass variant above a tile's breaks need the tile's global prefix
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
struct StreamWs {
  u32* pend; int* ph;          // provisional entries, max_pages * SS_PAGE each
  uint2* page_meta;            // page -> (warp, sequence number inside the warp's run)
  u32* page_ctr;               // pages handed out
  uint2* warp_tot;             // per warp: (sum of its deltas, its #breaks)
  ulonglong2* warp_base;       // per warp: (height, rank) at the start of its run
  uint4* marks;                // per chromosome: (warp, height, rank) at its first cell, run-relative
  u32 max_pages;
};

template <int NSTAGE>
__global__ void __launch_bounds__(512, 2)
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
  int4* wring = sm_x + w * (NSTAGE * 128);
  unsigned char* wlist = sm_list + w * 128;
  const int4* src = reinterpret_cast<const int4*>(delta) + (u64)s0 * 128 + lane;
  auto issue = [&](bool on, int stage, const int4* from) {
    if (on) {
      int4* dst = wring + stage * 128 + lane;
#pragma unroll
      for (int r = 0; r < 4; r++) cp_async16(dst + r * 32, from + r * 32);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < NSTAGE; st++) issue(s0 + st < s1, st, src + (u64)st * 128);

  // output pages: `cur` is being filled (sequence cur_seq), `nxt` is in hand so that a batch
  // may run over the page end; the page after that is requested as soon as `nxt` becomes `cur`
  u32 cur = 0, nxt = 0, cur_seq = 0, pend_reg = 0;
  bool pending = false;
  if (lane == 0) {
    cur = atomicAdd(W.page_ctr, 2u);
    if (cur + 1 < W.max_pages) { W.page_meta[cur] = make_uint2(gw, 0); W.page_meta[cur + 1] = make_uint2(gw, 1); }
  }
  cur = __shfl_sync(GR_FULL, cur, 0);
  nxt = cur + 1;
  auto resolve = [&]() {
    if (pending) {
      nxt = __shfl_sync(GR_FULL, pend_reg, 0);
      if (lane == 0 && nxt < W.max_pages) W.page_meta[nxt] = make_uint2(gw, cur_seq + 1);
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
    if (pg < W.max_pages) {
      const u64 a = ((u64)pg << SS_PAGE_SHIFT) | (idx & (SS_PAGE - 1));
      W.pend[a] = pos; W.ph[a] = (int)h;
    }
  };

  const u32 lt_mask = (1u << lane) - 1;
  u32 run_s = 0, run_c = 0;                            // height / #breaks since the start of the run
  u32 cur_blk = 0xffffffffu, jb_blk = 0, len = 0;
  int c = 0;
  bool act = false;
  u32 it = 0;
  for (u32 sp = s0; sp < s1; sp++, it++) {
    const u32 blk = sp >> 4;
    if (blk != cur_blk) {                              // chromosome of this 8192-cell block (warp-uniform)
      cur_blk = blk;
      c = L.blk2chrom[blk];
      jb_blk = (u32)(((u64)blk << GR_BLOCK_SHIFT) - L.off[c]);
      len = L.len[c];
      act = (L.flags[c] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);
    }
    const u32 jb = jb_blk + (sp & 15u) * 512u;         // chromosome position of the span's first cell
    if (jb == 0 && lane == 0) W.marks[c] = make_uint4(gw, run_s, run_c, 1u);
    cp_async_wait<NSTAGE - 1>();
    __syncwarp();
    const int4* wst = wring + (it % NSTAGE) * 128;
    int4* gcell = reinterpret_cast<int4*>(delta) + (u64)sp * 128;
    const bool fast = act && jb >= 1 && (u64)jb + 512 <= (u64)len;
    if (fast) {
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
      if (lane < 16) sm_bm[w * 16 + lane] = 0;
      __syncwarp();
      for (u32 base = 0; base < nnz; base += 32) {
        resolve();
        const u32 n = base + lane;
        const bool on = n < nnz;
        int4 x = make_int4(0, 0, 0, 0);
        u32 q = 0;
        if (on) { q = wlist[n]; x = wst[q]; }
        const u32 m4 = (x.x != 0 ? 1u : 0u) | (x.y != 0 ? 2u : 0u) | (x.z != 0 ? 4u : 0u) | (x.w != 0 ? 8u : 0u);
        const u32 c4 = __popc(m4);
        const u32 s1_ = (u32)x.x, s2_ = s1_ + (u32)x.y, s3_ = s2_ + (u32)x.z, s4_ = s3_ + (u32)x.w;
        const u32 inc_s = warp_incl_scan_u32(s4_, lane), inc_c = warp_incl_scan_u32(c4, lane);
        if (on) {
          const u32 h0 = run_s + inc_s - s4_;          // height before this chunk
          u32 idx = run_c + inc_c - c4;
          const u32 p0 = jb + q * 4;
          if (m4 & 1u) put(idx++, p0, h0);
          if (m4 & 2u) put(idx++, p0 + 1, h0 + s1_);
          if (m4 & 4u) put(idx++, p0 + 2, h0 + s2_);
          if (m4 & 8u) put(idx++, p0 + 3, h0 + s3_);
          atomicOr(&sm_bm[w * 16 + (q >> 3)], m4 << ((q & 7) * 4));
          // every break of an interior span is a non-zero cell and vice versa: clearing the
          // chunk leaves the delta array all zero for the next sample (no 4 B/bp memset)
          if (zero_after) gcell[q] = make_int4(0, 0, 0, 0);
        }
        run_s += __shfl_sync(GR_FULL, inc_s, 31);
        run_c += __shfl_sync(GR_FULL, inc_c, 31);
        if ((run_c >> SS_PAGE_SHIFT) > cur_seq) advance();
      }
      __syncwarp();
      if (lane < 16) bitmap[(u64)sp * 16 + lane] = sm_bm[w * 16 + lane];
    } else {
      // dense: chromosome ends, inactive chromosomes.  16 consecutive cells per lane.
      int d[SC_ITEMS];
      sc_load_items(wst, lane, d);
      u32 run = 0, m = 0;
      const u32 j0 = jb + lane * SC_ITEMS;
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        run += (u32)d[i];
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
    __syncwarp();                                      // every lane is done with the stage: refill it
    issue(sp + NSTAGE < s1, it % NSTAGE, src + (u64)(it + NSTAGE) * 128);
  }
  resolve();                                           // a page still on order gets its (unused) label
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
__global__ void __launch_bounds__(SS_PAGE)
k_scan_place(StreamWs W, DevRle out, int* __restrict__ err) {
  __shared__ float4 sm_lut[120];
  units_lut_fill(sm_lut, threadIdx.x, SS_PAGE);
  __syncthreads();
  const u32 npages = min(*W.page_ctr, W.max_pages);
  bool neg = false;
  for (u32 p = blockIdx.x; p < npages; p += gridDim.x) {
    const uint2 meta = W.page_meta[p];
    const u32 tot = W.warp_tot[meta.x].y, first = meta.y << SS_PAGE_SHIFT;
    if (first >= tot) continue;                        // the page a warp held in reserve
    const u32 n = min((u32)SS_PAGE, tot - first);
    if (threadIdx.x < n) {
      const ulonglong2 wb = W.warp_base[meta.x];
      const u64 a = ((u64)p << SS_PAGE_SHIFT) + threadIdx.x;
      const int N = (int)((u32)wb.x + (u32)W.ph[a]);
      neg |= N < 0;
      const u64 rank = wb.y + first + threadIdx.x;
      out.end[rank] = W.pend[a];
      out.val[rank] = units_to_val_lut(sm_lut, N < 0 ? 0 : N);
    }
  }
  if (neg) atomicOr(err, GR_DE_PILE);                  // ERRPILE 1921, 1969
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

template <int CT, int LAG, int CAP>
static void launch_dense_scan_t(cudaStream_t s, const DevLayout& L, int32_t* delta,
                                const ScanScratch& sc, DevRle out, u32* bitmap, int* err, int zero_after) {
  const u64 ntiles = L.T / (CT * 16);
  static int grid = 0;
  const size_t smem = (size_t)SC_NSTAGE * (CT * 4) * sizeof(int4) + (size_t)(LAG + 1) * CAP * sizeof(int2);
  auto kern = k_dense_scan<CT, LAG, CAP>;
  constexpr int SC_THREADS = CT + 32;
  if (!grid) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, SC_THREADS, smem);
    if (per < 1) per = 1;
    if (per > (CT == 512 ? 2 : 4)) per = CT == 512 ? 2 : 4;
    grid = sms * per;                                  // persistent: co-resident CTAs only
    if (grid > 1024) grid = 1024;                      // one lane per CTA group in the exchange
    if (getenv("GR_SCAN_DEBUG")) fprintf(stderr, "k_dense_scan<%d,%d,%d>: %d CTAs/SM, grid %d, smem %zu\n", CT, LAG, CAP, per, grid, smem);
  }
  unsigned g = (unsigned)(ntiles < (u64)grid ? ntiles : (u64)grid);
  const u64 nrounds = (ntiles + g - 1) / g;
  ScanStatus st;
  st.ngroups = (g + 31) / 32;
  st.agg = (ulonglong2*)sc.st_sum;
  st.grp = st.agg + ntiles;
  // sc.st_sum holds ntiles + nrounds*ngroups status words (allocated by the context)
  cudaMemsetAsync(sc.st_sum, 0, (ntiles + nrounds * st.ngroups) * sizeof(ulonglong2), s);
  u32 nt = (u32)ntiles;
  DevLayout Lc = L;
  void* args[] = { (void*)&delta, (void*)&Lc, (void*)&st, (void*)&out, (void*)&bitmap, (void*)&err, (void*)&nt, (void*)&zero_after };
  // cooperative launch: fails instead of deadlocking if the CTAs cannot all be resident
  cudaLaunchCooperativeKernel((const void*)kern, dim3(g), dim3(SC_THREADS), args, smem, s);
  GR_NOTE_LAUNCH();
  launch_fill_chrom_start(s, L, out.chrom_start, out.total);
}

size_t dense_scan_ws_bytes(u64 cap, int nchrom) {
  const u64 max_pages = (cap / SS_PAGE + 2 * SS_MAX_WARPS + 3) & ~1ull;
  return (size_t)(max_pages * SS_PAGE * 8 + max_pages * 8 + SS_MAX_WARPS * (8 + 16) + (u64)nchrom * 16 + 256);
}

static void launch_scan_stream(cudaStream_t s, const DevLayout& L, int32_t* delta,
                               const ScanScratch& sc, DevRle out, u32* bitmap, int* err, int zero_after) {
  static int grid = 0, nstage = 0;
  if (!grid) {
    const char* e = getenv("GR_SCAN_STAGES");
    nstage = e ? atoi(e) : 3;
    if (nstage != 2 && nstage != 4) nstage = 3;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    grid = sms * 2;
    if (grid * 16 > SS_MAX_WARPS) grid = SS_MAX_WARPS / 16;
    cudaFuncSetAttribute(k_scan_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 2048);
    cudaFuncSetAttribute(k_scan_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 3 * 2048);
    cudaFuncSetAttribute(k_scan_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4 * 2048);
  }
  StreamWs W;
  const u64 max_pages = (sc.cap / SS_PAGE + 2 * SS_MAX_WARPS + 3) & ~1ull;   // even: keeps the 16-byte arrays aligned
  char* p = (char*)sc.ws;
  W.pend = (u32*)p; p += max_pages * SS_PAGE * 4;
  W.ph = (int*)p; p += max_pages * SS_PAGE * 4;
  W.page_meta = (uint2*)p; p += max_pages * 8;
  W.warp_base = (ulonglong2*)p; p += SS_MAX_WARPS * 16;
  W.warp_tot = (uint2*)p; p += SS_MAX_WARPS * 8;
  W.marks = (uint4*)p; p += (u64)L.nchrom * 16;
  W.page_ctr = (u32*)p;
  W.max_pages = (u32)max_pages;
  cudaMemsetAsync(W.page_ctr, 0, 4, s);
  const u32 nspans = (u32)(L.T / 512);
  const u32 nwarps = (u32)grid * 16;
  const u32 R = (nspans + nwarps - 1) / nwarps;
  const size_t smem = (size_t)16 * nstage * 2048;
  if (nstage == 2) k_scan_stream<2><<<grid, 512, smem, s>>>(delta, L, W, bitmap, err, nspans, R, zero_after);
  else if (nstage == 4) k_scan_stream<4><<<grid, 512, smem, s>>>(delta, L, W, bitmap, err, nspans, R, zero_after);
  else k_scan_stream<3><<<grid, 512, smem, s>>>(delta, L, W, bitmap, err, nspans, R, zero_after);
  GR_NOTE_LAUNCH();
  k_scan_fix<<<1, 1024, 0, s>>>(L, W, out, err, nwarps); GR_NOTE_LAUNCH();
  k_scan_place<<<148 * 8, SS_PAGE, 0, s>>>(W, out, err); GR_NOTE_LAUNCH();
}

template <int LAG, int CAP, int NX>
static void launch_dense_scan_w(cudaStream_t s, const DevLayout& L, int32_t* delta,
                                const ScanScratch& sc, DevRle out, u32* bitmap, int* err, int zero_after) {
  const u64 ntiles = L.nblocks;                        // one tile per 8192-cell block
  static int grid = 0;
  const size_t smem = (size_t)16 * 2 * 128 * sizeof(int4) + (size_t)(2 * LAG) * CAP * sizeof(int2);
  auto kern = k_dense_scan_w<LAG, CAP, NX>;
  constexpr int threads = 512 + 32 * NX;
  if (!grid) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, threads, smem);
    if (per < 1) per = 1;
    if (per > 2) per = 2;
    grid = sms * per;                                  // persistent: co-resident CTAs only
    if (grid > 1024) grid = 1024;                      // one lane per CTA group in the exchange
    if (getenv("GR_SCAN_DEBUG")) fprintf(stderr, "k_dense_scan_w<%d,%d,%d>: %d CTAs/SM, grid %d, smem %zu\n", LAG, CAP, NX, per, grid, smem);
  }
  unsigned g = (unsigned)(ntiles < (u64)grid ? ntiles : (u64)grid);
  const u64 nrounds = (ntiles + g - 1) / g;
  ScanStatus st;
  st.ngroups = (g + 31) / 32;
  st.agg = (ulonglong2*)sc.st_sum;
  st.grp = st.agg + ntiles;
  cudaMemsetAsync(sc.st_sum, 0, (ntiles + nrounds * st.ngroups) * sizeof(ulonglong2), s);
  u32 nt = (u32)ntiles;
  DevLayout Lc = L;
  void* args[] = { (void*)&delta, (void*)&Lc, (void*)&st, (void*)&out, (void*)&bitmap, (void*)&err, (void*)&nt, (void*)&zero_after };
  cudaLaunchCooperativeKernel((const void*)kern, dim3(g), dim3(threads), args, smem, s);
  GR_NOTE_LAUNCH();
#ifdef GR_SCAN_PROF
  {
    static int nl = 0;
    if (++nl == 12) {
      unsigned long long h[8];
      cudaStreamSynchronize(s);
      cudaMemcpyFromSymbol(h, g_scan_prof, sizeof(h));
      const double n = (double)h[4], m = (double)h[6];
      fprintf(stderr, "scan prof (12 launches): per exchange: agg-wait %.0f ns, poll %.0f ns (%.2f polls), sibling %.0f ns; "
              "warp0 prefix wait %.0f ns per tile (%.0f exchanges, %.0f finishes)\n",
              h[0] / n, h[1] / n, h[3] / n, h[2] / n, h[5] / m, n, m);
    }
  }
#endif
  launch_fill_chrom_start(s, L, out.chrom_start, out.total);
}

void launch_dense_scan(cudaStream_t s, const DevLayout& L, int32_t* delta,
                       const ScanScratch& sc, DevRle out, u32* bitmap, int* err, int zero_after) {
  static int ver = -1;
  if (ver < 0) {
    const char* e = getenv("GR_SCAN_V");
    ver = e ? atoi(e) : 4;
  }
  if (ver == 4) { launch_scan_stream(s, L, delta, sc, out, bitmap, err, zero_after); return; }
  if (ver == 3) {
    static int wl = -1;
    if (wl < 0) { const char* e = getenv("GR_SCAN_LAG"); wl = e ? atoi(e) : 2; }
    static int nx = -1;
    if (nx < 0) { const char* e = getenv("GR_SCAN_NX"); nx = e ? atoi(e) : 2; }
    if (wl == 3) launch_dense_scan_w<3, 864, 2>(s, L, delta, sc, out, bitmap, err, zero_after);
    else if (nx == 1) launch_dense_scan_w<2, 1216, 1>(s, L, delta, sc, out, bitmap, err, zero_after);
    else if (nx == 3) launch_dense_scan_w<2, 1216, 3>(s, L, delta, sc, out, bitmap, err, zero_after);
    else launch_dense_scan_w<2, 1216, 2>(s, L, delta, sc, out, bitmap, err, zero_after);
    return;
  }
  static int lag = -1;
  if (lag < 0) {
    const char* e = getenv("GR_SCAN_LAG");             // tuning knob; default chosen from measurements
    lag = e ? atoi(e) : 2;
  }
  static int ct = -1;
  if (ct < 0) {
    const char* e = getenv("GR_SCAN_CT");
    ct = e ? atoi(e) : 512;
  }
  if (ct == 256) {
    if (lag == 3) launch_dense_scan_t<256, 3, 768>(s, L, delta, sc, out, bitmap, err, zero_after);
    else launch_dense_scan_t<256, 2, 768>(s, L, delta, sc, out, bitmap, err, zero_after);
  } else if (lag == 3) launch_dense_scan_t<512, 3, 1280>(s, L, delta, sc, out, bitmap, err, zero_after);
  else if (lag == 1) launch_dense_scan_t<512, 1, 1536>(s, L, delta, sc, out, bitmap, err, zero_after);
  else launch_dense_scan_t<512, 2, 1536>(s, L, delta, sc, out, bitmap, err, zero_after);
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
  // each CTA owns one contiguous slice of the interval array, so its running
  // chromosome changes at most a handful of times: sums stay in registers and
  // reach the per-chromosome counters with O(#CTAs) atomics instead of O(n/256)
  const u64 per = ((n + gridDim.x - 1) / gridDim.x + 1023) / 1024 * 1024;
  const u64 lo = (u64)blockIdx.x * per;
  const u64 hi = min(lo + per, n);
  u64 pi = 0, pf = 0;
  int cur = -1;                                        // chromosome the register sums belong to
  auto flush = [&]() {
    // block-wide: add (pi, pf) of all threads into chromosome `cur`
    u64 a = warp_sum_u64(pi), b = warp_sum_u64(pf);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sm_i[w] = a; sm_f[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0 && cur >= 0) {
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
    __syncthreads();
    if (threadIdx.x == 0) {
      sm_c0 = chrom_of_index(r.chrom_start, nchrom, base);
      sm_c1 = chrom_of_index(r.chrom_start, nchrom, last);
    }
    __syncthreads();
    const int c0 = sm_c0, c1 = sm_c1;
    if (c0 == c1) {
      if (c0 != cur) { flush(); cur = c0; }
      const u64 cs = r.chrom_start[c0];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const u64 i = base + j * 256 + threadIdx.x;
        if (i < hi) {
          const u32 e = r.end[i];
          const u32 st = (i == cs) ? 0u : r.end[i - 1];
          const float p = __fmul_rn(__uint2float_rn(e - st), r.val[i]);
          const u64 ip = (u64)p;                       // p >= 0
          pi += ip;
          pf += (u64)(__fsub_rn(p, (float)ip) * 1099511627776.0f);   // exact: fraction * 2^40
        }
      }
      if (pf >> 62) { pi += pf >> 40; pf &= (1ull << 40) - 1; }
    } else {
      // a chromosome boundary inside the round (rare): per-interval atomics
      flush();
      cur = -1;
      for (int j = 0; j < 4; j++) {
        const u64 i = base + j * 256 + threadIdx.x;
        if (i < hi) {
          const int c = chrom_of_index(r.chrom_start, nchrom, i);
          const u32 e = r.end[i];
          const u32 st = (i == r.chrom_start[c]) ? 0u : r.end[i - 1];
          const float p = __fmul_rn(__uint2float_rn(e - st), r.val[i]);
          const u64 ip = (u64)p;
          const u64 fp = (u64)(__fsub_rn(p, (float)ip) * 1099511627776.0f);
          if (ip) atomicAdd(acc_int + c, ip);
          if (fp) atomicAdd(acc_frac + c, fp);
        }
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
