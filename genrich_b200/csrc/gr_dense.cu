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
// K2: one pass over the dense int32 delta cells -- persistent, software-pipelined,
// warp-specialised.
//
//   grid = co-resident CTAs only (2 per SM, cooperative launch); tiles of 8192
//   cells are dealt round-robin: in round k CTA b works on tile kG+b.
//   Per CTA: 16 compute warps + 1 exchange warp.
//   * loads: cp.async (LDGSTS, 16 B/thread/op, coalesced) into a 3-stage XOR-swizzled
//     shared-memory ring; 1-2 tiles (32-64 KB) per CTA are always in flight to HBM.
//   * A(k)  compute warps: 16 consecutive cells per thread (conflict-free LDS.128),
//           thread/warp/block scan of (sum, #breaks), break bitmap written, tile
//           aggregate handed to the exchange warp.
//   * X(k)  exchange warp, concurrently with A(k+1): publishes the aggregate and
//           gathers the exclusive prefix of the tile with ONE L2 round trip:
//             agg[tile]      the tile's own (sum, #breaks)
//             grp[k*NG + g]  total of the 32 tiles of CTA-group g in round k
//           exclusive(kG+b) = (totals of all rounds < k, kept in registers)
//                           + sum_{g' < g} grp[k,g'] + sum_{b' in group, b' < b} agg.
//           Every dependency is on aggregates of the same or the previous round;
//           nothing waits for another tile's *prefix*, so there is no serial chain.
//   * B(k)  compute warps, after A(k+1): re-read tile k from its stage, add the
//           prefix, write each break as (end, value) at its global rank.
//   History (profiles/README.md): 32-wide decoupled look-back 0.92 TB/s; ticketed
//   persistent tiles 0.16 TB/s; 320-wide look-back 0.77 TB/s; two-level exchange
//   without the A/B split 1.03 TB/s (60 % of warp samples parked on the barrier
//   behind the exchange round trip).
//
// A break closes an interval at chromosome position j iff 1 <= j < len and
// delta[j] != 0, or j == len (Genrich.c:2241, 2268); its value is the running sum
// BEFORE delta[j] is added (2245), rebuilt as the reference float.
// Running sums are kept modulo 2^32: every true prefix fits in int32.
#define SC_CT 512                 // compute threads
#define SC_WARPS 16
#define SC_THREADS 544            // + one exchange warp
#define SC_ITEMS 16
#define SC_STAGE_INT4 2048        // 32 KB per stage
#define SC_NSTAGE 3
#define BAR_COMPUTE 1
#define BAR_AGG 2                 // +parity
#define BAR_PREFIX 4              // +parity

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

// 16-byte chunk g of a tile lives at chunk g ^ ((g >> 3) & 3) of its stage: the
// striped cp.async writes stay contiguous and the blocked reads (thread t reads
// chunks 4t..4t+3) hit 8 distinct 16-byte bank groups per quarter warp.
__device__ __forceinline__ int sc_swz(int g) { return g ^ ((g >> 3) & 3); }

__device__ __forceinline__ void sc_load_items(const int4* stage, int tid, int (&d)[SC_ITEMS]) {
  const int4 x0 = stage[sc_swz(4 * tid + 0)], x1 = stage[sc_swz(4 * tid + 1)];
  const int4 x2 = stage[sc_swz(4 * tid + 2)], x3 = stage[sc_swz(4 * tid + 3)];
  d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w;
  d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
  d[8] = x2.x; d[9] = x2.y; d[10] = x2.z; d[11] = x2.w;
  d[12] = x3.x; d[13] = x3.y; d[14] = x3.z; d[15] = x3.w;
}

__global__ void __launch_bounds__(SC_THREADS, 2)
k_dense_scan(const int32_t* __restrict__ delta, DevLayout L, ScanStatus S, DevRle out,
             u32* __restrict__ bitmap, int* __restrict__ err, u32 ntiles) {
  extern __shared__ int4 sm_x[];                       // SC_NSTAGE * SC_STAGE_INT4
  __shared__ u32 sm_wsum[SC_WARPS], sm_wcnt[SC_WARPS];
  __shared__ u32 sm_agg_sum[2], sm_agg_cnt[2];
  __shared__ u32 sm_ex_sum[2];
  __shared__ u64 sm_ex_cnt[2];

  const int tid = threadIdx.x, lane = tid & 31;
  const u32 G = gridDim.x, b = blockIdx.x;

  // ------------------------------------------------------------ exchange warp
  if (tid >= SC_CT) {
    const u32 g = b >> 5, j = b & 31, ng = S.ngroups;
    u32 rnd_s = 0;                                     // totals of all rounds before k
    u64 rnd_c = 0;
    u32 k = 0;
    for (u32 tile = b; tile < ntiles; tile += G, k++) {
      named_sync(BAR_AGG + (k & 1), 64);               // aggregate of tile k is in shared memory
      const u32 agg_s = sm_agg_sum[k & 1], agg_c = sm_agg_cnt[k & 1];
      if (lane == 0) st_status(S.agg + tile, 1, agg_s, agg_c);
      const u32 last_b = min(G - 1, ntiles - 1 - k * G);
      const bool need_a = (u32)lane < j, need_g = (u32)lane < g, need_p = k > 0 && (u32)lane < ng;
      const ulonglong2* pa = S.agg + (tile - j) + lane;
      const ulonglong2* pg = S.grp + (u64)k * ng + lane;
      const ulonglong2* pp = S.grp + (u64)(k - 1) * ng + lane;   // only dereferenced when k > 0
      // (1) the group's own aggregates: as soon as they are in, the group's last tile
      //     publishes the group total -- it must NOT wait for the totals of earlier
      //     groups, or the ten groups of a round serialise (measured: 35 polls/tile)
      ulonglong2 va;
      for (;;) {
        va.x = va.y = 0;
        if (need_a) va = ld_status(pa);
        const bool ok = !need_a || ((va.x >> 62) == 1 && (va.y >> 62) == 1);
        if (__all_sync(GR_FULL, ok)) break;
      }
      u32 s_in = need_a ? (u32)va.x : 0u;
      u64 c_in = need_a ? (va.y & GR_LB_PAYLOAD) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_in += __shfl_xor_sync(GR_FULL, s_in, o);
        c_in += __shfl_xor_sync(GR_FULL, c_in, o);
      }
      if (lane == 0 && b == min(32u * g + 31u, last_b))
        st_status(S.grp + (u64)k * ng + g, 1, s_in + agg_s, c_in + agg_c);
      // (2) totals of the earlier groups of this round and of all groups of the previous round
      ulonglong2 vg, vp;
      for (;;) {
        vg.x = vg.y = vp.x = vp.y = 0;
        if (need_g) vg = ld_status(pg);
        if (need_p) vp = ld_status(pp);
        const bool ok = (!need_g || ((vg.x >> 62) == 1 && (vg.y >> 62) == 1)) &&
                        (!need_p || ((vp.x >> 62) == 1 && (vp.y >> 62) == 1));
        if (__all_sync(GR_FULL, ok)) break;
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
        sm_ex_sum[k & 1] = rnd_s + s_g + s_in;
        sm_ex_cnt[k & 1] = rnd_c + c_g + c_in;
      }
      __syncwarp();
      named_arrive(BAR_PREFIX + (k & 1), SC_THREADS);  // prefix of tile k is in shared memory
    }
    return;
  }

  // ------------------------------------------------------------ compute warps
  const int w = tid >> 5;
  auto issue = [&](u32 tile, int stage) {              // chunk g (16 B) -> swizzled slot
    if (tile < ntiles) {
      const int4* src = reinterpret_cast<const int4*>(delta + (u64)tile * GR_BLOCK_SLOTS);
      int4* dst = sm_x + stage * SC_STAGE_INT4;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int gch = q * SC_CT + tid;
        cp_async16(dst + sc_swz(gch), src + gch);
      }
    }
    cp_async_commit();
  };

  issue(b, 0);
  issue(b + G, 1);
  TileMeta meta = tile_meta(L, b, ntiles);

  // state of the tile whose breaks are still to be written (B phase)
  u32 p_tile = 0, p_jb = 0, p_pre_sum = 0, p_wx_cnt = 0, p_lane_cnt_incl = 0, p_m = 0, p_tcnt = 0;
  int p_c = 0;
  bool p_first = false, have_prev = false;

  auto emit_prev = [&](u32 k_prev) {
    named_sync(BAR_PREFIX + (k_prev & 1), SC_THREADS);
    const u32 ex_sum = sm_ex_sum[k_prev & 1];
    const u64 ex_cnt = sm_ex_cnt[k_prev & 1];
    if (tid == 0) {
      if (p_first) {
        out.chrom_start[p_c] = ex_cnt;
        if (ex_sum != 0) atomicOr(err, GR_DE_TAIL);    // previous chromosome did not return to 0 (2283-2289)
      }
      if (p_tile == ntiles - 1) {
        *out.total = ex_cnt + p_tcnt;
        out.chrom_start[L.nchrom] = ex_cnt + p_tcnt;
      }
    }
    // Breaks are ~6 % of the cells: walking 16 predicated per-item blocks with one or
    // two live lanes each made the kernel instruction-bound (ncu r1d: 42 thread
    // instructions per cell, 19 of 32 lanes active).  Instead every lane drops its
    // (position, height) pairs into the warp's own 2 KB slice of the stage it has
    // just read (cheap, sparse), and the warp then converts and stores them densely:
    // lane n handles the n-th break, so the global stores are fully coalesced.
    const u32 wcnt = __shfl_sync(GR_FULL, p_lane_cnt_incl, 31);          // breaks of this warp
    if (wcnt) {                                                          // warp-uniform
      int4* stage4 = sm_x + (k_prev % SC_NSTAGE) * SC_STAGE_INT4;
      int d[SC_ITEMS];
      sc_load_items(stage4, tid, d);
      __syncwarp();                                                      // every lane has its cells
      const u64 wrank = ex_cnt + p_wx_cnt;                               // rank of the warp's first break
      const u32 wpos0 = p_jb + (u32)w * (32 * SC_ITEMS);                 // chromosome position of the warp's first cell
      u32 run = ex_sum + p_pre_sum;                                      // exclusive prefix before d[0]
      if (wcnt <= 256) {
        int2* st = reinterpret_cast<int2*>(stage4 + w * 128);            // 256 entries of 8 B
        u32 r = p_lane_cnt_incl - __popc(p_m);
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
          if (p_m & (1u << i)) st[r++] = make_int2(lane * SC_ITEMS + i, (int)run);
          run += (u32)d[i];
        }
        __syncwarp();
        bool neg = false;
        for (u32 n = lane; n < wcnt; n += 32) {
          const int2 e = st[n];
          neg |= e.y < 0;
          out.end[wrank + n] = wpos0 + (u32)e.x;
          out.val[wrank + n] = units_to_val(e.y < 0 ? 0 : e.y);
        }
        if (neg) atomicOr(err, GR_DE_PILE);                              // ERRPILE 1921, 1969
      } else if (p_m) {                                                  // > 256 breaks in 512 cells: direct path
        u64 rank = wrank + (p_lane_cnt_incl - __popc(p_m));
        bool neg = false;
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
          if (p_m & (1u << i)) {
            const int N = (int)run;
            neg |= N < 0;
            out.end[rank] = wpos0 + lane * SC_ITEMS + i;
            out.val[rank] = units_to_val(N < 0 ? 0 : N);
            rank++;
          }
          run += (u32)d[i];
        }
        if (neg) atomicOr(err, GR_DE_PILE);
      }
    }
  };

  u32 k = 0;
  for (u32 tile = b; tile < ntiles; tile += G, k++) {
    // ---- A(k): scan tile k, publish its aggregate
    // chromosome of the tile after this one: first hop now, second hop after the scan
    const int c_next = tile + G < ntiles ? L.blk2chrom[tile + G] : 0;
    cp_async_wait<1>();
    named_sync(BAR_COMPUTE, SC_CT);                    // tile k is in shared memory (all threads' copies)
    int d[SC_ITEMS];
    sc_load_items(sm_x + (k % SC_NSTAGE) * SC_STAGE_INT4, tid, d);
    const u64 tbase = (u64)tile * GR_BLOCK_SLOTS;
    const u32 jb = (u32)(tbase - meta.off);            // chromosome position of the tile's first cell
    const u32 len = meta.len;
    const bool interior = jb >= 1 && (u64)jb + GR_BLOCK_SLOTS <= (u64)len;
    u32 run = 0, m = 0;
    if (interior) {
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        run += (u32)d[i];
        m |= (d[i] != 0 ? 1u : 0u) << i;
      }
    } else {
      const u32 j0 = jb + tid * SC_ITEMS;
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        run += (u32)d[i];
        const u32 jj = j0 + i;
        const bool brk = (jj == len) || (d[i] != 0 && jj >= 1 && jj < len);
        m |= (brk ? 1u : 0u) << i;
      }
    }
    if (!meta.act) m = 0;
    const u32 cnt = __popc(m);
    const u32 wi_sum = warp_incl_scan_u32(run, lane);
    const u32 wi_cnt = warp_incl_scan_u32(cnt, lane);
    if (lane == 31) { sm_wsum[w] = wi_sum; sm_wcnt[w] = wi_cnt; }
    {
      const u32 hi = __shfl_down_sync(GR_FULL, m, 1);
      if (!(lane & 1)) bitmap[(tbase >> 5) + (tid >> 1)] = m | (hi << 16);
    }
    named_sync(BAR_COMPUTE, SC_CT);
    u32 wx_sum = 0, wx_cnt = 0, t_sum = 0, t_cnt = 0;
#pragma unroll
    for (int q = 0; q < SC_WARPS; q++) {
      const u32 a = sm_wsum[q], c2 = sm_wcnt[q];
      if (q < w) { wx_sum += a; wx_cnt += c2; }
      t_sum += a; t_cnt += c2;
    }
    if (w == 0) {
      if (lane == 0) { sm_agg_sum[k & 1] = t_sum; sm_agg_cnt[k & 1] = t_cnt; }
      __syncwarp();
      named_arrive(BAR_AGG + (k & 1), 64);             // hand over to the exchange warp
    }
    const u32 n_pre_sum = wx_sum + (wi_sum - run);
    TileMeta meta_next;
    meta_next.c = c_next;
    meta_next.off = L.off[c_next];
    meta_next.len = L.len[c_next];
    meta_next.act = (L.flags[c_next] & (GR_CF_OWNED | GR_CF_SAVE)) == (GR_CF_OWNED | GR_CF_SAVE);

    // ---- B(k-1): write the breaks of the previous tile (its prefix arrived meanwhile)
    if (have_prev) emit_prev(k - 1);
    named_sync(BAR_COMPUTE, SC_CT);                    // stage (k+2)%3 == (k-1)%3 is free; sm_w* reusable
    issue(tile + 2 * G, (k + 2) % SC_NSTAGE);

    p_tile = tile; p_jb = jb; p_pre_sum = n_pre_sum; p_wx_cnt = wx_cnt; p_lane_cnt_incl = wi_cnt; p_m = m; p_tcnt = t_cnt;
    p_c = meta.c; p_first = tbase == meta.off; have_prev = true;
    meta = meta_next;
  }
  if (have_prev) emit_prev(k - 1);
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
  const size_t smem = (size_t)SC_NSTAGE * SC_STAGE_INT4 * sizeof(int4);
  if (!grid) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_dense_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_dense_scan, SC_THREADS, smem);
    if (per < 1) per = 1;
    if (per > 2) per = 2;
    grid = sms * per;                                  // persistent: co-resident CTAs only
    if (grid > 1024) grid = 1024;                      // one lane per CTA group in the exchange
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
  void* args[] = { (void*)&delta, (void*)&Lc, (void*)&st, (void*)&out, (void*)&bitmap, (void*)&err, (void*)&nt };
  // cooperative launch: fails instead of deadlocking if the CTAs cannot all be resident
  cudaLaunchCooperativeKernel((const void*)k_dense_scan, dim3(g), dim3(SC_THREADS), args, smem, s);
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
  // each CTA owns one contiguous slice of the interval array, so its running
  // chromosome changes at most a handful of times: sums stay in registers and
  // reach the per-chromosome counters with O(#CTAs) atomics instead of O(n/256)
  const u64 per = ((n + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
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
  for (u64 base = lo; base < hi; base += 256) {
    const u64 last = min(base + 256, hi) - 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      sm_c0 = chrom_of_index(r.chrom_start, nchrom, base);
      sm_c1 = chrom_of_index(r.chrom_start, nchrom, last);
    }
    __syncthreads();
    const int c0 = sm_c0, c1 = sm_c1;
    const u64 i = base + threadIdx.x;
    if (c0 == c1) {
      if (c0 != cur) { flush(); cur = c0; }
      if (i < hi) {
        const u32 e = r.end[i];
        const u32 st = (i == r.chrom_start[c0]) ? 0u : r.end[i - 1];
        const float p = __fmul_rn(__uint2float_rn(e - st), r.val[i]);
        const u64 ip = (u64)p;                         // p >= 0
        pi += ip;
        pf += (u64)(__fsub_rn(p, (float)ip) * 1099511627776.0f);   // exact: fraction * 2^40
        if (pf >> 62) { pi += pf >> 40; pf &= (1ull << 40) - 1; }
      }
    } else {
      // a chromosome boundary inside the tile (rare): per-interval atomics
      flush();
      cur = -1;
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
