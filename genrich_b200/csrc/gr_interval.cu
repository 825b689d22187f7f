// gr_interval.cu -- interval-domain kernels between the dense scan and the peak
// scan: K3 control sweep, K4 breakpoint union, K5 -log10 p through a table of
// distinct (expt, ctrl) pairs, K6 Fisher combine.
#include <stdlib.h>
#include <string.h>
#include "gr_tile.cuh"
#include "gr_math.cuh"
#include "gr_internal.h"

// ============================================================================
// K3: control sweep (savePileupCtrl 2103-2141).  Per raw control interval:
// net = MAX(factor * val, lambda) in float; an interval boundary survives iff the
// clamped value changes across it (2122) or it is the chromosome end.  Surviving
// intervals are compacted (single pass, look-back ranks); the bits of the dropped
// boundaries are cleared in the control break bitmap.
#define CL_TILE 8192          // raw intervals per look-back tile (8 warps x 32 rounds x 32 lanes)

__device__ __forceinline__ float clamp_net(float factor, float v, float lambda) {
  if (v == -1.0f) return -1.0f;                    // SKIP: inside a -E region (2124); never equals a clamped value
  const float s = __fmul_rn(factor, v);
  return s > lambda ? s : lambda;                 // MAX(val, lambda), Genrich.h:12
}

__global__ void __launch_bounds__(256)
k_ctrl_clamp(DevLayout L, DevRle raw, const float* __restrict__ fl, Lookback<1> lb,
             DevRle out, u32* __restrict__ bitmap) {
  // launched for the capacity of the raw array; the interval count and the two scalars
  // (scale factor, lambda) are read from device memory: no host round trip in between
  const u64 n = *raw.total;
  if ((u64)blockIdx.x * CL_TILE >= n) return;            // tickets stay dense
  const float factor = fl[0], lambda = fl[1];
  const u32 tile = take_ticket(lb.ticket);
  const u64 t0 = (u64)tile * CL_TILE;
  const u64 tl = min(t0 + CL_TILE, n) - 1;
  const TileChrom tc = tile_chrom_range(raw.chrom_start, L.nchrom, t0, tl);
  const bool uni = tc.c0 == tc.c1;
  const u64 uni_end = uni ? raw.chrom_start[tc.c0 + 1] : 0;      // one past the chromosome's last interval
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u64 wbase = t0 + (u64)w * 1024;

  // pass 1: which boundaries survive; dropped ones leave the break bitmap right away
  u32 mine = 0, cnt = 0;
#pragma unroll 4
  for (int k = 0; k < 32; k++) {
    const u64 i = wbase + k * 32 + lane;
    bool keep = false;
    if (i < n) {
      const float net = clamp_net(factor, raw.val[i], lambda);
      int c = tc.c0;
      bool last;
      if (uni) last = i + 1 == uni_end;
      else { c = chrom_of_index(raw.chrom_start, L.nchrom, i); last = i + 1 == raw.chrom_start[c + 1]; }
      keep = last || net != clamp_net(factor, raw.val[i + 1], lambda);
      if (!keep) {
        const u64 g = L.off[c] + raw.end[i];
        atomicAnd(bitmap + (g >> 5), ~(1u << (g & 31)));
      }
    }
    const u32 bal = __ballot_sync(GR_FULL, keep);
    if (lane == k) mine = bal;
    cnt += __popc(bal);
  }
  u32 tot;
  u64 r = tile_exclusive_rank(lb, tile, lane == 31 ? cnt : 0u, tot);
  r = __shfl_sync(GR_FULL, r, 31);
  if (tile == 0 && threadIdx.x == 0) out.chrom_start[0] = 0;

  // pass 2: write the survivors at their rank (values re-read: still in L1/L2)
  for (int k = 0; k < 32; k++) {
    const u32 bal = __shfl_sync(GR_FULL, mine, k);
    if (bal & (1u << lane)) {
      const u64 i = wbase + k * 32 + lane;
      const u64 rank = r + __popc(bal & ((1u << lane) - 1));
      out.end[rank] = raw.end[i];
      out.val[rank] = clamp_net(factor, raw.val[i], lambda);
      int c = tc.c0;
      bool last;
      if (uni) last = i + 1 == uni_end;
      else { c = chrom_of_index(raw.chrom_start, L.nchrom, i); last = i + 1 == raw.chrom_start[c + 1]; }
      if (last) out.chrom_start[c + 1] = rank + 1;
      if (i == n - 1) *out.total = rank + 1;
    }
    r += __popc(bal);
  }
}

// chromosomes without intervals take the running count (forward fill)
__global__ void k_fill_forward(int nchrom, const u64* raw_start, u64* out_start) {
  if (threadIdx.x || blockIdx.x) return;
  for (int c = 0; c < nchrom; c++)
    if (raw_start[c + 1] == raw_start[c]) out_start[c + 1] = out_start[c];
}

void launch_ctrl_clamp(cudaStream_t s, const DevLayout& L, const DevRle& raw, u64 n_upper,
                       const float* factor_lambda, const CompactScratch& sc,
                       DevRle out, u32* bitmap) {
  cudaMemsetAsync(out.total, 0, sizeof(u64), s);
  if (!n_upper) return;
  const u64 ntiles = (n_upper + CL_TILE - 1) / CL_TILE;
  cudaMemsetAsync(sc.st, 0, ntiles * sizeof(u64), s);
  cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s);
  Lookback<1> lb;
  lb.st[0] = sc.st; lb.ticket = sc.ticket;
  k_ctrl_clamp<<<(unsigned)ntiles, 256, 0, s>>>(L, raw, factor_lambda, lb, out, bitmap); GR_NOTE_LAUNCH();
  k_fill_forward<<<1, 32, 0, s>>>(L.nchrom, raw.chrom_start, out.chrom_start); GR_NOTE_LAUNCH();
}

// no control: lambda over every active chromosome, SKIP inside its -E regions (saveLambda
// 1838-1877).  The host knows the partition (chromosome ends and region boundaries) and writes
// the ends; the values (lambda lives on the device) and the break bits are set here.
__global__ void k_ctrl_const(const float* __restrict__ lambda, u64 n, float* __restrict__ val, u32* __restrict__ bitmap,
                             const u64* __restrict__ slots, const uint8_t* __restrict__ skip) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  val[i] = skip[i] ? -1.0f : *lambda;
  const u64 g = slots[i];
  atomicOr(bitmap + (g >> 5), 1u << (g & 31));
}
void launch_ctrl_const(cudaStream_t s, const DevLayout& L, const float* lambda_dev, u64 n, DevRle out, u32* bitmap,
                       const u64* slots, const uint8_t* skip) {
  cudaMemsetAsync(bitmap, 0, (L.T / 32) * sizeof(u32), s);
  if (n) { k_ctrl_const<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(lambda_dev, n, out.val, bitmap, slots, skip); GR_NOTE_LAUNCH(); }
}

// fixed-point per-chromosome sums (integer part, 2^-40 fraction) -> doubles
__global__ void k_sums_double(const u64* __restrict__ acc_int, const u64* __restrict__ acc_frac, int nchrom,
                              double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nchrom) out[c] = (double)acc_int[c] + (double)acc_frac[c] * (1.0 / 1099511627776.0);
}
void launch_sums_double(cudaStream_t s, const u64* acc_int, const u64* acc_frac, int nchrom, double* out) {
  k_sums_double<<<(nchrom + 127) / 128, 128, 0, s>>>(acc_int, acc_frac, nchrom, out); GR_NOTE_LAUNCH();
}

// calcLambda 1817-1832 / calcFactor 2043-2045 on the device: the per-chromosome doubles are
// added in chromosome order by one thread, like the reference's running sums
__global__ void k_lambda_factor(const double* __restrict__ sums, int nchrom, int has_ctrl, u64 genome_len,
                                float* __restrict__ factor_lambda, int* __restrict__ err) {
  if (threadIdx.x || blockIdx.x) return;
  double f = 0.0, g = 0.0;
  for (int c = 0; c < nchrom; c++) { f += sums[c]; if (has_ctrl) g += sums[nchrom + c]; }
  if (f == 0.0) atomicOr(err, GR_DE_EXPT);                       // Genrich.c:2292
  float factor = 1.0f;
  if (has_ctrl && g != 0.0) factor = (float)(f / g);
  factor_lambda[0] = factor;
  factor_lambda[1] = (float)(f / (double)genome_len);
}
void launch_lambda_factor(cudaStream_t s, const double* sums, int nchrom, bool has_ctrl, u64 genome_len,
                          float* factor_lambda, int* err) {
  k_lambda_factor<<<1, 32, 0, s>>>(sums, nchrom, has_ctrl ? 1 : 0, genome_len, factor_lambda, err); GR_NOTE_LAUNCH();
}

// ============================================================================
// K4: union of experimental and control breakpoints (savePval 1768-1791,
// countIntervals 1661) on the break bitmaps.  Pass A ranks the set bits of E, C and
// E|C per 8192-cell block (look-back over three counters); pass B walks the set
// bits of E|C and gathers the pileup values: the interval ending at a break lies
// in the experimental interval number (#E breaks before it), same for control.
#define UR_BLOCKS 16        // bitmap blocks (of 256 words) per look-back tile
__global__ void __launch_bounds__(256)
k_union_rank(const u32* __restrict__ bmE, const u32* __restrict__ bmC, Lookback<3> lb,
             u64* __restrict__ rankE, u64* __restrict__ rankC, u64* __restrict__ rankU,
             u64* __restrict__ totals, u32 nblocks, u32 ntiles) {
  __shared__ u32 sm_blk[UR_BLOCKS][3];
  __shared__ i64 sm_ex[3];
  const u32 tile = take_ticket(lb.ticket);
  const u32 b0 = tile * UR_BLOCKS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // a warp takes UR_BLOCKS / 8 whole blocks (8 words per lane and bitmap of each): the popcounts add up in
  // registers and one packed warp reduction per block is left (a reduction per word and counter -- 48 per
  // thread -- was what this kernel spent its time on: 0.5 ms for 0.77 GB)
  constexpr int PER = UR_BLOCKS / 8;
  u32 E[PER][8], C[PER][8];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const u32 b = b0 + w * PER + q;
    const bool on = b < nblocks;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const u64 widx = (u64)b * 256 + i * 32 + lane;
      E[q][i] = on ? bmE[widx] : 0u;
      C[q][i] = (on && bmC) ? bmC[widx] : 0u;
    }
  }
#pragma unroll
  for (int q = 0; q < PER; q++) {
    u32 ac = 0, u = 0;                                 // <= 8192 per block: E and C counts travel packed
#pragma unroll
    for (int i = 0; i < 8; i++) {
      ac += __popc(E[q][i]) | (__popc(C[q][i]) << 16);
      u += __popc(E[q][i] | C[q][i]);
    }
    ac = __reduce_add_sync(GR_FULL, ac);
    u = __reduce_add_sync(GR_FULL, u);
    if (lane == 0) { sm_blk[w * PER + q][0] = ac & 0xffffu; sm_blk[w * PER + q][1] = ac >> 16; sm_blk[w * PER + q][2] = u; }
  }
  __syncthreads();
  if (w == 0) {
    i64 agg[3] = { 0, 0, 0 }, ex[3];
    for (int b = 0; b < UR_BLOCKS; b++) { agg[0] += sm_blk[b][0]; agg[1] += sm_blk[b][1]; agg[2] += sm_blk[b][2]; }
    lookback_exclusive<3, 2>(lb, tile, agg, ex);
    if (lane == 0) {
      sm_ex[0] = ex[0]; sm_ex[1] = ex[1]; sm_ex[2] = ex[2];
      if (tile == ntiles - 1) {
        totals[0] = (u64)(ex[0] + agg[0]);
        totals[1] = (u64)(ex[1] + agg[1]);
        totals[2] = (u64)(ex[2] + agg[2]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < UR_BLOCKS && b0 + threadIdx.x < nblocks) {
    u64 e = (u64)sm_ex[0], c = (u64)sm_ex[1], u = (u64)sm_ex[2];
    for (u32 b = 0; b < threadIdx.x; b++) { e += sm_blk[b][0]; c += sm_blk[b][1]; u += sm_blk[b][2]; }
    rankE[b0 + threadIdx.x] = e;
    if (rankC) rankC[b0 + threadIdx.x] = c;
    rankU[b0 + threadIdx.x] = u;
  }
}

void launch_union_rank(cudaStream_t s, const DevLayout& L, const u32* bmE, const u32* bmC,
                       const RankScratch& sc, u64* rankE, u64* rankC, u64* rankU, u64* totals) {
  const u32 ntiles = (u32)((L.nblocks + UR_BLOCKS - 1) / UR_BLOCKS);
  for (int k = 0; k < 3; k++) cudaMemsetAsync(sc.st[k], 0, (size_t)ntiles * sizeof(u64), s);
  cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s);
  Lookback<3> lb;
  for (int k = 0; k < 3; k++) lb.st[k] = sc.st[k];
  lb.ticket = sc.ticket;
  k_union_rank<<<ntiles, 256, 0, s>>>(bmE, bmC, lb, rankE, rankC, rankU, totals, (u32)L.nblocks, ntiles); GR_NOTE_LAUNCH();
}

// Pass B.  A CTA takes UE_BLOCKS consecutive bitmap blocks.  Walking the set bits thread by
// thread (first version) made every warp wait for its fullest word -- five or six dependent
// DRAM round trips for gathers that were two cache lines wide -- and kept only two loads per
// thread in flight (1.34 ms per hg38 replicate, 0.4 of the HBM rate).  Here the breaks are first
// listed in shared memory in rank order (cell offset, experimental / control interval number
// relative to the CTA's first block); the list is then streamed by all threads, UE_UNROLL
// entries per thread at a time: gathers and stores are coalesced and every thread has
// 2 * UE_UNROLL independent loads in flight.
#define UE_CAP 4096            // list entries per round (a CTA with more breaks takes several rounds)
#define UE_UNROLL 4
template <int UE_BLOCKS>
__global__ void __launch_bounds__(256)
k_union_emit(DevLayout L, const u32* __restrict__ bmE, const u32* __restrict__ bmC,
             const u64* __restrict__ rankE, const u64* __restrict__ rankC,
             const u64* __restrict__ rankU, const float* __restrict__ exptVal,
             const float* __restrict__ ctrlVal, u32* __restrict__ pEnd,
             float* __restrict__ pExpt, float* __restrict__ pCtrl, u32* __restrict__ bmU,
             u64* __restrict__ chrom_start, u32 nblocks) {
  __shared__ u32 sm_ent[UE_CAP];                 // bits 0-14: cell offset inside the CTA's blocks, 15-31: experimental interval number
  __shared__ unsigned short sm_ctl[UE_CAP];      // control interval number
  __shared__ u32 sm_a[UE_BLOCKS * 8], sm_u[UE_BLOCKS * 8];
  __shared__ u32 sm_jb[UE_BLOCKS];
  __shared__ u32 sm_tot;
  const u32 b0 = blockIdx.x * UE_BLOCKS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 E[UE_BLOCKS], C[UE_BLOCKS];
#pragma unroll
  for (int k = 0; k < UE_BLOCKS; k++) {
    const bool on = b0 + k < nblocks;
    const u64 widx = (u64)(b0 + k) * 256 + threadIdx.x;
    E[k] = on ? bmE[widx] : 0u;
    C[k] = on ? bmC[widx] : 0u;
  }
  const u64 RE0 = rankE[b0], RC0 = rankC[b0], RU0 = rankU[b0];
  if (threadIdx.x < UE_BLOCKS && b0 + threadIdx.x < nblocks) {
    const u32 blk = b0 + threadIdx.x;
    const int c = L.blk2chrom[blk];
    const u64 off = L.off[c];
    sm_jb[threadIdx.x] = (u32)((u64)blk * GR_BLOCK_SLOTS - off);
    if ((u64)blk * GR_BLOCK_SLOTS == off) chrom_start[c] = rankU[blk];
  }
  // exclusive ranks of every word inside the CTA's range, order (block, word): E and C counts
  // travel packed (<= 32768 each), U alone
  u32 xa[UE_BLOCKS], xu[UE_BLOCKS];
#pragma unroll
  for (int k = 0; k < UE_BLOCKS; k++) {
    const u32 U = E[k] | C[k];
    if (b0 + k < nblocks) bmU[(u64)(b0 + k) * 256 + threadIdx.x] = U;
    const u32 pa = __popc(E[k]) | (__popc(C[k]) << 16), pu = __popc(U);
    const u32 ia = warp_incl_scan_u32(pa, lane), iu = warp_incl_scan_u32(pu, lane);
    if (lane == 31) { sm_a[k * 8 + w] = ia; sm_u[k * 8 + w] = iu; }
    xa[k] = ia - pa; xu[k] = iu - pu;
  }
  __syncthreads();
  if (w == 0) {                                  // UE_BLOCKS * 8 <= 32 warp totals -> exclusive
    const u32 va = lane < UE_BLOCKS * 8 ? sm_a[lane] : 0u, vu = lane < UE_BLOCKS * 8 ? sm_u[lane] : 0u;
    const u32 ia = warp_incl_scan_u32(va, lane), iu = warp_incl_scan_u32(vu, lane);
    if (lane < UE_BLOCKS * 8) { sm_a[lane] = ia - va; sm_u[lane] = iu - vu; }
    if (lane == 31) sm_tot = iu;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < UE_BLOCKS; k++) { xa[k] += sm_a[k * 8 + w]; xu[k] += sm_u[k * 8 + w]; }
  const u32 tot = sm_tot;
  for (u32 lo = 0; lo < tot; lo += UE_CAP) {
    if (lo) __syncthreads();                     // the previous round's list has been streamed
#pragma unroll
    for (int k = 0; k < UE_BLOCKS; k++) {
      u32 U = E[k] | C[k];
      u32 u = xu[k];
      if (u >= lo + UE_CAP || u + __popc(U) <= lo) continue;
      const u32 cell0 = ((u32)k << GR_BLOCK_SHIFT) | (threadIdx.x << 5);
      while (U) {
        const int b = __ffs(U) - 1;
        const u32 low = (1u << b) - 1;
        if (u >= lo && u < lo + UE_CAP) {
          sm_ent[u - lo] = (cell0 + b) | (((xa[k] & 0xffff) + __popc(E[k] & low)) << 15);
          sm_ctl[u - lo] = (unsigned short)((xa[k] >> 16) + __popc(C[k] & low));
        }
        u++;
        U &= U - 1;
      }
    }
    __syncthreads();
    const u32 cnt = min(tot - lo, (u32)UE_CAP);
    for (u32 i0 = threadIdx.x; i0 < cnt; i0 += 256 * UE_UNROLL) {
      u32 en[UE_UNROLL];
      float ve[UE_UNROLL], vc[UE_UNROLL];
#pragma unroll
      for (int q = 0; q < UE_UNROLL; q++) {
        const u32 i = i0 + q * 256;
        if (i < cnt) {
          en[q] = sm_ent[i];
          ve[q] = exptVal[RE0 + (en[q] >> 15)];
          vc[q] = ctrlVal[RC0 + sm_ctl[i]];
        }
      }
#pragma unroll
      for (int q = 0; q < UE_UNROLL; q++) {
        const u32 i = i0 + q * 256;
        if (i < cnt) {
          const u64 u = RU0 + lo + i;
          pEnd[u] = sm_jb[(en[q] >> GR_BLOCK_SHIFT) & (UE_BLOCKS - 1)] + (en[q] & (GR_BLOCK_SLOTS - 1));
          pExpt[u] = ve[q];
          pCtrl[u] = vc[q];
        }
      }
    }
  }
}

void launch_union_emit(cudaStream_t s, const DevLayout& L, const u32* bmE, const u32* bmC,
                       const u64* rankE, const u64* rankC, const u64* rankU,
                       const float* exptVal, const float* ctrlVal,
                       u32* pEnd, float* pExpt, float* pCtrl, u32* bmU, u64* chrom_start,
                       const u64* total) {
  const unsigned grid = (unsigned)((L.nblocks + 3) / 4);       // four bitmap blocks per CTA
  k_union_emit<4><<<grid, 256, 0, s>>>(L, bmE, bmC, rankE, rankC, rankU, exptVal, ctrlVal,
                                       pEnd, pExpt, pCtrl, bmU, chrom_start, (u32)L.nblocks);
  GR_NOTE_LAUNCH();
  launch_fill_chrom_start(s, L, chrom_start, total);
}

// ============================================================================
// K5: -log10 p per merged interval (calcPval 1628).  p is a pure function of the
// (expt, ctrl) float pair and the pairs take few distinct values, so the FP64
// evaluation runs once per distinct pair: open-addressing table keyed by the 64
// bits of the pair, then a gather.  The same table code, keyed by the 32 bits of
// -log10 p and accumulating interval lengths, is the genome-wide histogram of
// hashPval 300 / recordPval 277 (exact float equality == equal bit patterns here:
// all keys are >= +0).
#define TBL_EMPTY (~0ull)

__device__ __forceinline__ u32 mix64(u64 k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (u32)k;
}

// find-or-insert; returns slot, or ~0u when the probe sequence is exhausted
__device__ __forceinline__ u32 table_upsert(const PairTable& t, u64 key, bool& fresh) {
  const u32 mask = t.cap - 1;
  u32 h = mix64(key) & mask;
  fresh = false;
  for (u32 probe = 0; probe < 512; probe++) {
    u64 k = t.keys[h];
    if (k == key) return h;
    if (k == TBL_EMPTY) {
      k = atomicCAS(t.keys + h, TBL_EMPTY, key);
      if (k == TBL_EMPTY) { fresh = true; return h; }
      if (k == key) return h;
    }
    h = (h + 1) & mask;
  }
  return ~0u;
}

// The first probe of PI_UNROLL keys is in flight together (the table is L2-resident, a probe
// is one L2 round trip; one key at a time left the kernel waiting on it).
#define PI_UNROLL 4
__global__ void __launch_bounds__(256)
k_pair_insert(const float* __restrict__ pExpt, const float* __restrict__ pCtrl, const u64* __restrict__ n_dev,
              PairTable t, u32* __restrict__ slot, int* __restrict__ err) {
  const u64 n = *n_dev;                                  // interval count, device side
  const u64 stride = (u64)gridDim.x * (256 * PI_UNROLL);
  const u32 mask = t.cap - 1;
  bool bad = false;
  for (u64 i0 = (u64)blockIdx.x * (256 * PI_UNROLL); i0 < n; i0 += stride) {    // warp-uniform trip count
    u64 key[PI_UNROLL], k0[PI_UNROLL];
    u32 h[PI_UNROLL];
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++) {
      const u64 i = i0 + q * 256 + threadIdx.x;
      key[q] = i < n ? ((u64)__float_as_uint(pExpt[i]) << 32) | __float_as_uint(pCtrl[i]) : 0ull;
    }
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++) {
      h[q] = mix64(key[q]) & mask;
      k0[q] = t.keys[h[q]];
    }
    u32 nf = 0;
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++) {
      const u64 i = i0 + q * 256 + threadIdx.x;
      bool fresh = false;
      if (i < n) {
        const u32 hs = k0[q] == key[q] ? h[q] : table_upsert(t, key[q], fresh);
        if (hs == ~0u) bad = true;
        slot[i] = hs == ~0u ? 0u : hs;
      }
      nf += __popc(__ballot_sync(GR_FULL, fresh));
    }
    if ((threadIdx.x & 31) == 0 && nf) {
      const u32 tot = atomicAdd(t.count, nf) + nf;
      if (tot > (t.cap >> 1)) bad = true;
    }
  }
  if (bad) atomicOr(err, GR_DE_TABLE);
}

static unsigned capped_grid(u64 n_upper) {
  const u64 b = (n_upper + 255) / 256;
  return (unsigned)(b < 148 * 32 ? (b ? b : 1) : 148 * 32);
}
void launch_pair_insert(cudaStream_t s, const float* pExpt, const float* pCtrl, u64 n_upper, const u64* n_dev,
                        const PairTable& t, u32* slot, int* err) {
  if (!n_upper) return;
  k_pair_insert<<<capped_grid(n_upper), 256, 0, s>>>(pExpt, pCtrl, n_dev, t, slot, err); GR_NOTE_LAUNCH();
}

__global__ void __launch_bounds__(128)
k_pair_eval(PairTable t) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.cap) return;
  const u64 k = t.keys[i];
  if (k == TBL_EMPTY) return;
  t.pval[i] = gm_calc_pval(__uint_as_float((u32)(k >> 32)), __uint_as_float((u32)k));
}
void launch_pair_eval(cudaStream_t s, const PairTable& t) {
  k_pair_eval<<<(t.cap + 127) / 128, 128, 0, s>>>(t); GR_NOTE_LAUNCH();
}

__global__ void __launch_bounds__(256)
k_gather_f32(const float* __restrict__ table, const u32* __restrict__ slot, const u64* __restrict__ n_dev,
             float* __restrict__ out) {
  const u64 n = *n_dev;
  const u64 stride = (u64)gridDim.x * (256 * PI_UNROLL);
  for (u64 i0 = (u64)blockIdx.x * (256 * PI_UNROLL) + threadIdx.x; i0 < n; i0 += stride) {
    u32 sl[PI_UNROLL];
    float v[PI_UNROLL];
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++) sl[q] = i0 + q * 256 < n ? slot[i0 + q * 256] : ~0u;
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++) v[q] = sl[q] == ~0u ? -1.0f : table[sl[q]];   // ~0: SKIP interval, never entered in the table
#pragma unroll
    for (int q = 0; q < PI_UNROLL; q++)
      if (i0 + q * 256 < n) out[i0 + q * 256] = v[q];
  }
}
void launch_gather_f32(cudaStream_t s, const float* table, const u32* slot, u64 n_upper, const u64* n_dev, float* out) {
  if (n_upper) { k_gather_f32<<<capped_grid(n_upper), 256, 0, s>>>(table, slot, n_dev, out); GR_NOTE_LAUNCH(); }
}

// histogram insert keyed by the bits of -log10 p; lengths are added with one
// atomic per distinct key per warp (hot keys such as p == 0 would otherwise
// serialise in L2).  SKIP (-1) is not recorded (hashPval 319).
__global__ void __launch_bounds__(256)
k_key_insert(const u32* __restrict__ pEnd, const float* __restrict__ pval, u64 n,
             const u64* __restrict__ chrom_start, int nchrom, PairTable t,
             u32* __restrict__ slot, int* __restrict__ err) {
  const u64 t0 = (u64)blockIdx.x * blockDim.x;
  const u64 tl = min(t0 + blockDim.x, n) - 1;
  const TileChrom tc = tile_chrom_range(chrom_start, nchrom, t0, tl);
  const u64 i = t0 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool fresh = false, bad = false;
  u32 h = 0xfffffffeu;                 // lanes without a key never match a real slot
  u32 len = 0;
  if (i < n) {
    const float p = pval[i];
    if (p != -1.0f) {
      const int c = tc.c0 == tc.c1 ? tc.c0 : chrom_of_index(chrom_start, nchrom, i);
      const u32 e = pEnd[i];
      len = e - (i == chrom_start[c] ? 0u : pEnd[i - 1]);
      h = table_upsert(t, (u64)__float_as_uint(p), fresh);
      bad = h == ~0u;
      if (bad) h = 0xfffffffeu;
      slot[i] = bad ? 0u : h;
    } else
      slot[i] = ~0u;
  }
  // warp-aggregate equal slots
  const u32 peers = __match_any_sync(GR_FULL, h);
  u64 tot = 0;
  if (peers == (1u << lane)) tot = len;
  else {
    for (u32 rem = peers; rem; rem &= rem - 1) {
      // every lane of the group walks the same member list
      const int src = __ffs(rem) - 1;
      tot += __shfl_sync(peers, len, src);
    }
  }
  if (h < 0xfffffffeu && lane == __ffs(peers) - 1 && tot) atomicAdd(t.lens + h, tot);
  const u32 nf = __popc(__ballot_sync(GR_FULL, fresh));
  if (lane == 0 && nf) {
    const u32 c2 = atomicAdd(t.count, nf) + nf;
    if (c2 > (t.cap >> 1)) bad = true;
  }
  if (bad) atomicOr(err, GR_DE_TABLE);
}

void launch_key_insert(cudaStream_t s, const u32* pEnd, const float* pval, u64 n,
                       const u64* chrom_start, int nchrom, const PairTable& t, u32* slot, int* err) {
  if (!n) return;
  k_key_insert<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pEnd, pval, n, chrom_start, nchrom, t, slot, err); GR_NOTE_LAUNCH();
}

// The same histogram when the final p array is ONE replicate's (no Fisher combine): every interval already knows
// the slot of its (expt, ctrl) pair in the pair table (k_pair_insert) and the pair's -log10 p sits in t.pval, so
// the bp are added per SLOT -- no hashing, no probing, no second table.  Distinct pairs with the same p simply
// show up as repeated keys in the list; the BH pass merges equal keys anyway (it has to, for the all-gather).
// 8 intervals per thread; equal slots inside a warp are added up first (hot pairs such as (0, lambda)).
// Hot pairs -- (0, lambda) alone covers most of a genome -- would serialise in L2 even after the warp-level
// merge (224 M intervals of the ATAC configuration: 6.3 ms): every CTA first collects its sums in a direct-mapped
// (Tried and removed: four intervals per thread in flight and one warp round per distinct slot instead of
// __match_any_sync -- 1.5x slower on the ATAC sample and 5x slower on the multimapped 10 Gbp shard, whose warps
// hold many distinct slots.)
// shared-memory cache of SH_CACHE slots (tag = table slot; a slot that finds its cache line taken goes straight to
// global memory) and flushes the cache once at the end.
#define SH_CACHE 2048
__global__ void __launch_bounds__(256)
k_slot_hist(const u32* __restrict__ pEnd, const u32* __restrict__ slot, const u64* __restrict__ n_dev,
            u64* __restrict__ lens) {
  __shared__ u32 sm_tag[SH_CACHE];                   // table slot + 1; 0: free
  __shared__ u64 sm_val[SH_CACHE];
  for (int i = threadIdx.x; i < SH_CACHE; i += 256) { sm_tag[i] = 0; sm_val[i] = 0; }
  __syncthreads();
  const u64 n = *n_dev;
  const int lane = threadIdx.x & 31;
  const u64 stride = (u64)gridDim.x * 256;
  for (u64 i0 = (u64)blockIdx.x * 256; i0 < n; i0 += stride) {         // block-uniform trip count
    const u64 i = i0 + threadIdx.x;
    u32 h = 0xfffffffeu, len = 0;
    if (i < n) {
      h = slot[i];
      const u32 e = pEnd[i];
      u32 st = i ? pEnd[i - 1] : 0u;
      if (st >= e) st = 0u;                        // first interval of a chromosome: the previous end belongs to another one
      len = e - st;                                // (ends increase strictly inside a chromosome)
    }
    const u32 peers = __match_any_sync(GR_FULL, h);
    u64 tot = 0;
    if (peers == (1u << lane)) tot = len;
    else
      for (u32 rem = peers; rem; rem &= rem - 1) tot += __shfl_sync(peers, len, __ffs(rem) - 1);   // the group walks its member list
    if (h < 0xfffffffeu && lane == __ffs(peers) - 1 && tot) {
      const u32 ci = (h * 2654435761u) >> (32 - 11);             // SH_CACHE = 2^11 lines
      const u32 old = atomicCAS(&sm_tag[ci], 0u, h + 1u);
      if (old == 0u || old == h + 1u) atomicAdd(&sm_val[ci], tot);
      else atomicAdd(lens + h, tot);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SH_CACHE; i += 256)
    if (sm_tag[i] && sm_val[i]) atomicAdd(lens + (sm_tag[i] - 1u), sm_val[i]);
}
void launch_slot_hist(cudaStream_t s, const u32* pEnd, const u32* slot, u64 n_upper, const u64* n_dev,
                      const u64* chrom_start, int nchrom, u64* lens) {
  if (!n_upper) return;
  const u64 tiles = (n_upper + 255) / 256;
  k_slot_hist<<<(unsigned)(tiles < 148 * 4 ? tiles : 148 * 4), 256, 0, s>>>(pEnd, slot, n_dev, lens); GR_NOTE_LAUNCH();
  (void)chrom_start; (void)nchrom;
}

// table -> dense (key, len) list
// by_pval: the table is the PAIR table -- the list's key is the pair's -log10 p (t.pval), pairs that evaluate to
// SKIP (-E regions) or that no interval with bp refers to are left out
__global__ void __launch_bounds__(256)
k_table_compact(PairTable t, Lookback<1> lb, u32* __restrict__ keys_out,
                u64* __restrict__ lens_out, u64* __restrict__ count_out, u32 ntiles, int by_pval) {
  const u32 tile = take_ticket(lb.ticket);
  const u32 i = tile * 256 + threadIdx.x;
  u64 k = i < t.cap ? t.keys[i] : TBL_EMPTY;
  if (by_pval && k != TBL_EMPTY) {
    const float p = t.pval[i];
    k = p == -1.0f ? TBL_EMPTY : (u64)__float_as_uint(p);
  }
  const u32 f = k != TBL_EMPTY;
  u32 tot;
  const u64 r = tile_exclusive_rank(lb, tile, f, tot);
  if (f) { keys_out[r] = (u32)k; lens_out[r] = t.lens[i]; }
  if (tile == ntiles - 1 && threadIdx.x == 255) *count_out = r + f;
}

void launch_table_compact(cudaStream_t s, const PairTable& t, const CompactScratch& sc,
                          u32* keys_out, u64* lens_out, u64* count_out, int by_pval) {
  const u32 ntiles = (t.cap + 255) / 256;
  cudaMemsetAsync(sc.st, 0, (size_t)ntiles * sizeof(u64), s);
  cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s);
  Lookback<1> lb;
  lb.st[0] = sc.st; lb.ticket = sc.ticket;
  k_table_compact<<<ntiles, 256, 0, s>>>(t, lb, keys_out, lens_out, count_out, ntiles, by_pval); GR_NOTE_LAUNCH();
}

// q of every occupied slot: binary search of its key among the distinct keys
__global__ void __launch_bounds__(256)
k_table_q(PairTable t, const u32* __restrict__ dk, const float* __restrict__ dq,
          const u64* __restrict__ dcount, int by_pval) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.cap) return;
  const u64 k = t.keys[i];
  if (k == TBL_EMPTY) return;
  u32 key = (u32)k;
  if (by_pval) {                                   // pair table: look the pair's p up; a SKIP pair stays SKIP
    const float p = t.pval[i];
    if (p == -1.0f) { t.qval[i] = -1.0f; return; }
    key = __float_as_uint(p);
  }
  u64 lo = 0, hi = *dcount;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (dk[mid] < key) lo = mid + 1; else hi = mid;
  }
  t.qval[i] = dq[lo];
}
void launch_table_q(cudaStream_t s, const PairTable& t, const u32* dk, const float* dq, const u64* dcount, int by_pval) {
  k_table_q<<<(t.cap + 255) / 256, 256, 0, s>>>(t, dk, dq, dcount, by_pval); GR_NOTE_LAUNCH();
}

// ============================================================================
// K6: Fisher's method over replicates (combinePval 612-667, multPval 567-583).
// The union of the replicates' breakpoints is the OR of their union bitmaps; at
// each combined break the replicate's interval number is its own bit rank.
__global__ void __launch_bounds__(256)
k_or_bitmaps(const u32* __restrict__ a, const u32* __restrict__ b, u32* __restrict__ out, u64 n) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] | b[i];
}
void launch_or_bitmaps(cudaStream_t s, const u32* a, const u32* b, u32* out, u64 nwords) {
  if (nwords) { k_or_bitmaps<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(a, b, out, nwords); GR_NOTE_LAUNCH(); }
}

void launch_block_rank(cudaStream_t s, const DevLayout& L, const u32* bm, const CompactScratch& sc,
                       u64* rank, u64* total) {
  // reuse the three-counter kernel with C absent: totals[] needs 3 entries
  RankScratch rs;
  rs.st[0] = sc.st; rs.st[1] = sc.st + L.nblocks; rs.st[2] = sc.st + 2 * L.nblocks;
  rs.ticket = sc.ticket;
  launch_union_rank(s, L, bm, nullptr, rs, rank, nullptr, rank, total);
}

// One CTA per 8192-cell block, one thread per bitmap word, FE_CHUNK replicates per launch.  The replicates' array
// pointers arrive as kernel PARAMETERS and everything that depends only on the block (the combined word, every
// replicate's word and rank base) is requested at once; the p-values of an interval are then gathered together
// and its sum and df built in registers and written once.  (The first version walked the replicates one by one
// through a view table in global memory, with a read-modify-write of sum / df per replicate: five dependent
// round trips per CTA and 7.6 GB of DRAM traffic -- 9-10 ms per hg38 run of three replicates.)
// Summation order = replicate order (multPval 570-574), in double.
#define FE_CHUNK 8
struct FisherChunk {                                 // by value: the pointers are read from the parameter bank, not through memory
  const u32* bmU[FE_CHUNK]; const u64* rankU[FE_CHUNK]; const float* pval[FE_CHUNK]; const uint8_t* present[FE_CHUNK];
  int n;                                             // replicates in this chunk
  int first;                                         // 1: the chunk starts the sums (else they are continued)
};
__global__ void __launch_bounds__(256)
k_fisher_emit(DevLayout L, const u32* __restrict__ bmAll, const u64* __restrict__ rankAll, FisherChunk R,
              u32* __restrict__ end_out, double* __restrict__ sum_out, int* __restrict__ df_out,
              u64* __restrict__ chrom_start) {
  __shared__ u32 sm_w[FE_CHUNK + 1][8];
  const u32 blk = blockIdx.x;
  const u64 widx = (u64)blk * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // everything that depends on nothing but the block: one round trip
  const u32 A = bmAll[widx];
  const int c = L.blk2chrom[blk];
  const u64 rA = rankAll[blk];
  u32 Rw[FE_CHUNK];
  u64 rb[FE_CHUNK];
#pragma unroll
  for (int k = 0; k < FE_CHUNK; k++) {
    Rw[k] = k < R.n ? R.bmU[k][widx] : 0u;
    rb[k] = k < R.n ? R.rankU[k][blk] : 0ull;
  }
  // second hop: what hangs off the chromosome
  const u64 off = L.off[c];
  bool on[FE_CHUNK];
#pragma unroll
  for (int k = 0; k < FE_CHUNK; k++) on[k] = k < R.n && R.present[k][c];   // pval[j] == NULL (571): no such chromosome in the replicate
  const u32 pa = __popc(A);
  const u32 ia = warp_incl_scan_u32(pa, lane);
  if (lane == 31) sm_w[FE_CHUNK][w] = ia;
  u32 xr[FE_CHUNK];
#pragma unroll
  for (int k = 0; k < FE_CHUNK; k++) {
    if (!on[k]) Rw[k] = 0u;
    const u32 pr = __popc(Rw[k]);
    const u32 ir = warp_incl_scan_u32(pr, lane);
    if (lane == 31) sm_w[k][w] = ir;
    xr[k] = ir - pr;
  }
  __syncthreads();
  u32 xa = ia - pa;
  for (int j = 0; j < w; j++) xa += sm_w[FE_CHUNK][j];
#pragma unroll
  for (int k = 0; k < FE_CHUNK; k++) {
    u32 x = xr[k];
    for (int j = 0; j < w; j++) x += sm_w[k][j];
    rb[k] += x;
  }
  if (R.first && threadIdx.x == 0 && (u64)blk * GR_BLOCK_SLOTS == off) chrom_start[c] = rA;
  const u32 jw = (u32)((u64)blk * GR_BLOCK_SLOTS + (u64)threadIdx.x * 32 - off);
  u64 u = rA + xa;
  for (u32 m = A; m; m &= m - 1, u++) {
    const int b = __ffs(m) - 1;
    const u32 low = (1u << b) - 1;
    float p[FE_CHUNK];
#pragma unroll
    for (int k = 0; k < FE_CHUNK; k++) p[k] = on[k] ? R.pval[k][rb[k] + __popc(Rw[k] & low)] : -1.0f;   // the gathers, together
    double sum = 0.0;
    int df = 0;
    if (!R.first) { sum = sum_out[u]; df = df_out[u]; }
#pragma unroll
    for (int k = 0; k < FE_CHUNK; k++)
      if (p[k] != -1.0f) { sum += (double)p[k]; df += 2; }      // replicate order (570-574); SKIP is not counted (572)
    if (R.first) end_out[u] = jw + b;
    sum_out[u] = sum;
    df_out[u] = df;
  }
}

void launch_fisher_emit(cudaStream_t s, const DevLayout& L, const u32* bmAll, const u64* rankAll,
                        const RepView* reps_host, int nrep, u32* end_out, double* sum_out,
                        int* df_out, u64* chrom_start, const u64* total) {
  for (int r0 = 0; r0 < nrep; r0 += FE_CHUNK) {
    FisherChunk R;
    memset(&R, 0, sizeof R);
    R.n = nrep - r0 < FE_CHUNK ? nrep - r0 : FE_CHUNK;
    R.first = r0 == 0;
    for (int k = 0; k < R.n; k++) {
      R.bmU[k] = reps_host[r0 + k].bmU; R.rankU[k] = reps_host[r0 + k].rankU;
      R.pval[k] = reps_host[r0 + k].pval; R.present[k] = reps_host[r0 + k].present;
    }
    k_fisher_emit<<<(unsigned)L.nblocks, 256, 0, s>>>(L, bmAll, rankAll, R, end_out, sum_out, df_out, chrom_start);
    GR_NOTE_LAUNCH();
  }
  launch_fill_chrom_start(s, L, chrom_start, total);
}

// multPval 567-583 per combined interval.  The chi-square tail (R's pgamma series, FP64) is evaluated once per
// DISTINCT sum: the sums are sums of a few replicate p-values that each take few distinct values, so 150 M
// intervals of an hg38 run share a few hundred thousand sums (evaluating every interval took 22 ms).
//   k_fsum_insert  trivial cases at once (df 0: SKIP; df 2 or sum 0: the sum itself, 577); the rest find-or-insert
//                  the bits of their sum in the table (the creator of a slot leaves its df beside the key)
//   k_fsum_eval    one pchisq per occupied slot
//   k_fsum_gather  an interval whose df is the slot's takes the slot's value; the (never seen, but possible) interval
//                  with the same sum and ANOTHER df evaluates its own
__device__ __forceinline__ float fisher_p(double s, int d) {
  const double p = gm_pchisq(2.0 * s / GR_LOG10E, d);
  return p > (double)FLT_MAX ? FLT_MAX : (float)p;
}
__global__ void __launch_bounds__(256)
k_fsum_insert(const double* __restrict__ sum, const int* __restrict__ df, u64 n, PairTable t,
              u32* __restrict__ slot, float* __restrict__ out, int* __restrict__ err) {
  const u64 stride = (u64)gridDim.x * 256;
  bool bad = false;
  for (u64 i0 = (u64)blockIdx.x * 256; i0 < n; i0 += stride) {          // block-uniform trip count
    const u64 i = i0 + threadIdx.x;
    bool fresh = false;
    if (i < n) {
      const int d = df[i];
      const double s = sum[i];
      if (d == 0) { out[i] = -1.0f; slot[i] = ~0u; }
      else if (d == 2 || s == 0.0) { out[i] = (float)s; slot[i] = ~0u; }
      else {
        const u32 hs = table_upsert(t, (u64)__double_as_longlong(s), fresh);
        if (hs == ~0u) { bad = true; out[i] = fisher_p(s, d); slot[i] = ~0u; fresh = false; }
        else { slot[i] = hs; if (fresh) t.lens[hs] = (u64)d; }
      }
    }
    const u32 nf = __popc(__ballot_sync(GR_FULL, fresh));
    if ((threadIdx.x & 31) == 0 && nf) {
      const u32 tot = atomicAdd(t.count, nf) + nf;
      if (tot > (t.cap >> 1)) bad = true;
    }
  }
  if (bad) atomicOr(err, GR_DE_TABLE);
}
__global__ void __launch_bounds__(128)
k_fsum_eval(PairTable t) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.cap) return;
  const u64 k = t.keys[i];
  if (k == TBL_EMPTY) return;
  t.pval[i] = fisher_p(__longlong_as_double((long long)k), (int)t.lens[i]);
}
__global__ void __launch_bounds__(256)
k_fsum_gather(const double* __restrict__ sum, const int* __restrict__ df, u64 n, PairTable t,
              const u32* __restrict__ slot, float* __restrict__ out) {
  const u64 stride = (u64)gridDim.x * 256;
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const u32 sl = slot[i];
    if (sl == ~0u) continue;                                    // written by the insert pass
    const int d = df[i];
    out[i] = (int)t.lens[sl] == d ? t.pval[sl] : fisher_p(sum[i], d);
  }
}
void launch_fisher_table(cudaStream_t s, const double* sum, const int* df, u64 n, const PairTable& t, u32* slot,
                         float* pcomb, int* err) {
  if (!n) return;
  k_fsum_insert<<<capped_grid(n), 256, 0, s>>>(sum, df, n, t, slot, pcomb, err); GR_NOTE_LAUNCH();
  k_fsum_eval<<<(t.cap + 127) / 128, 128, 0, s>>>(t); GR_NOTE_LAUNCH();
  k_fsum_gather<<<capped_grid(n), 256, 0, s>>>(sum, df, n, t, slot, pcomb); GR_NOTE_LAUNCH();
}

__global__ void __launch_bounds__(128)
k_fisher_eval(const double* __restrict__ sum, const int* __restrict__ df, u64 n,
              float* __restrict__ out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = df[i];
  const double s = sum[i];
  float r;
  if (d == 0) r = -1.0f;
  else if (d == 2 || s == 0.0) r = (float)s;
  else r = fisher_p(s, d);
  out[i] = r;
}
void launch_fisher_eval(cudaStream_t s, const double* sum, const int* df, u64 n, float* pcomb) {
  if (n) { k_fisher_eval<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(sum, df, n, pcomb); GR_NOTE_LAUNCH(); }
}

