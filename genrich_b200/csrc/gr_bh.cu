// gr_bh.cu -- K7: Benjamini-Hochberg q-values over the genome-wide histogram of
// distinct -log10 p (computeQval 352, collectPval 333, saveQval 212-229).
//
// Input: (key = float bits of -log10 p, len = bp) pairs, possibly with repeated
// keys (several ranks' histograms concatenated by the all-gather).  Device radix
// sort by key (LSD, 4 x 8 bits, stable), merge equal keys, then from the largest
// key down:  k = 1 + sum of len over strictly larger keys (220, 228),
//   x = p + logN + log10f(k)   (two float adds, 226),   q = MAX(MIN(x, q_next), 0).
// Keys are >= +0 so the unsigned bit pattern orders exactly like the float.
#include "gr_tile.cuh"
#include "gr_internal.h"
#include <float.h>

#define RS_CHUNK 4096          // elements per block per pass
#define RS_THREADS 256

// log10f and logf as glibc 2.39 computes them -- saveQval 226 adds log10f(k) to p in float, so a last-bit difference
// shows in q (and in the sixth decimal of a -f line).  The reference's arithmetic lives in a third-party dependency
// that is not part of its sources (glibc 2.39 libm, x86-64); its published algorithms are restated here:
//   logf   (sysdeps/ieee754/flt-32/e_logf.c, the ARM optimized-routines logf): 16-entry table of (1/c, log c),
//          r = z/c - 1, degree-3 polynomial, all in double, rounded to float once.  Checked on the host against the
//          installed libm for ALL 2^23 floats in [1, 2) -- the only arguments log10f hands it -- with and without
//          fused multiply-adds: identical.
//   log10f (sysdeps/ieee754/flt-32/e_log10f.c): exponent and mantissa apart, y * log10_2lo + ivln10 * logf(m), then
//          + y * log10_2hi, plain float operations.  Checked against the installed libm over 9e7 counts: identical.
// x: a count of bp as a float, >= 1.
__device__ __forceinline__ float logf_glibc_1_2(float x) {      // x in [1, 2)
  const double T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5}, {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2}, {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
  const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  const u32 ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  const u32 tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
  const double r = __dadd_rn(__dmul_rn(z, T[i][0]), -1.0);
  const double y0 = __dadd_rn(T[i][1], __dmul_rn((double)k, Ln2));
  const double r2 = __dmul_rn(r, r);
  double y = __dadd_rn(__dmul_rn(A1, r), A2);
  y = __dadd_rn(__dmul_rn(A0, r2), y);
  y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
  return (float)y;
}
__device__ __forceinline__ float log10f_glibc(float x) {
  const float ivln10 = 4.3429449201e-01f, log10_2hi = 3.0102920532e-01f, log10_2lo = 7.9034151668e-07f;
  const int hx = __float_as_int(x);
  const int k = (hx >> 23) - 127;                      // >= 0 here
  const float y = (float)k;
  const float m = __int_as_float((hx & 0x007fffff) | (0x7f << 23));
  const float lg = logf_glibc_1_2(m);
  const float z = __fadd_rn(__fmul_rn(y, log10_2lo), __fmul_rn(ivln10, lg));
  return __fadd_rn(z, __fmul_rn(y, log10_2hi));
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const u32* __restrict__ keys, u64 n, int shift, u32* __restrict__ hist, u32 nblk) {
  __shared__ u32 h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * RS_CHUNK;
  for (int r = 0; r < RS_CHUNK / RS_THREADS; r++) {
    const u64 i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(u64)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];   // bin-major
}

// exclusive scan of the bin-major histogram (256 * nblk entries), one block
__global__ void __launch_bounds__(1024)
k_rs_scan(u32* __restrict__ hist, u64 m) {
  __shared__ u32 sm_w[32];
  __shared__ u32 sm_carry;
  if (threadIdx.x == 0) sm_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u64 base = 0; base < m; base += 1024) {
    const u64 i = base + threadIdx.x;
    const u32 v = i < m ? hist[i] : 0;
    const u32 wi = warp_incl_scan_u32(v, lane);
    if (lane == 31) sm_w[w] = wi;
    __syncthreads();
    u32 wx = 0, tot = 0;
    for (int k = 0; k < 32; k++) { const u32 a = sm_w[k]; if (k < w) wx += a; tot += a; }
    const u32 carry = sm_carry;
    if (i < m) hist[i] = carry + wx + wi - v;
    __syncthreads();
    if (threadIdx.x == 0) sm_carry = carry + tot;
    __syncthreads();
  }
}

// stable scatter: rounds of 256 elements in index order; rank inside a round =
// (same digit in lower warps) + (same digit in lower lanes of this warp)
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const u32* __restrict__ kin, const u64* __restrict__ lin, u64 n, int shift,
             const u32* __restrict__ hist, u32 nblk, u32* __restrict__ kout, u64* __restrict__ lout) {
  __shared__ u32 base[256];
  __shared__ u32 wc[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  base[threadIdx.x] = hist[(u64)threadIdx.x * nblk + blockIdx.x];
  for (int k = 0; k < 8; k++) wc[k][threadIdx.x] = 0;
  __syncthreads();
  const u64 cbase = (u64)blockIdx.x * RS_CHUNK;
  for (int r = 0; r < RS_CHUNK / RS_THREADS; r++) {
    const u64 i = cbase + r * RS_THREADS + threadIdx.x;
    const bool on = i < n;
    const u32 key = on ? kin[i] : 0;
    const u32 d = on ? ((key >> shift) & 255) : 0x100u + lane;   // inactive lanes never match
    const u32 peers = __match_any_sync(GR_FULL, d);
    const u32 rank_w = __popc(peers & ((1u << lane) - 1));
    const bool leader = on && (lane == __ffs(peers) - 1);
    if (leader) wc[w][d] = __popc(peers);
    __syncthreads();
    u32 lower = 0, total = 0;
    if (on) {
#pragma unroll
      for (int k = 0; k < 8; k++) { const u32 a = wc[k][d]; if (k < w) lower += a; total += a; }
      const u32 pos = base[d] + lower + rank_w;
      kout[pos] = key;
      lout[pos] = lin[i];
    }
    __syncthreads();
    if (leader) {
      if (lower == 0) base[d] += total;     // lowest warp holding this digit advances the base
      wc[w][d] = 0;
    }
    __syncthreads();
  }
}

// heads of equal-key runs -> distinct keys; lengths of a run summed into its slot
__global__ void __launch_bounds__(256)
k_bh_distinct(const u32* __restrict__ k, const u64* __restrict__ l, u64 n, Lookback<1> lb,
              u32* __restrict__ dk, u64* __restrict__ dl, u64* __restrict__ dcount, u32 ntiles) {
  const u32 tile = take_ticket(lb.ticket);
  const u64 i = (u64)tile * 256 + threadIdx.x;
  const bool on = i < n;
  const u32 key = on ? k[i] : 0;
  const u32 head = on && (i == 0 || k[i - 1] != key);
  u32 tot;
  const u64 r = tile_exclusive_rank(lb, tile, head, tot);     // heads before this element
  if (on) {
    const u64 slot = r + head - 1;                            // run index of element i
    if (head) dk[slot] = key;
    atomicAdd(dl + slot, l[i]);
    if (i == n - 1) *dcount = slot + 1;
  }
}

// descending sweep over the distinct keys, one block, chunks of 1024 from the top
__global__ void __launch_bounds__(1024)
k_bh_q(const u32* __restrict__ dk, const u64* __restrict__ dl, const u64* __restrict__ dcount,
       float logN, float* __restrict__ dq) {
  __shared__ u64 sm_s[32];
  __shared__ float sm_m[32];
  __shared__ u64 sm_carry_k;
  __shared__ float sm_carry_q;
  const u64 D = *dcount;
  if (threadIdx.x == 0) { sm_carry_k = 1; sm_carry_q = FLT_MAX; }   // k = 1 (220), qVal[pLen] = FLT_MAX (223)
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (u64 done = 0; done < D; done += 1024) {
    const u64 pos = done + threadIdx.x;          // 0 = largest key
    const bool on = pos < D;
    const u64 i = on ? D - 1 - pos : 0;
    const u64 len = on ? dl[i] : 0;
    // exclusive prefix (over larger keys) of len
    u64 inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(GR_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm_s[w] = inc;
    __syncthreads();
    u64 wx = 0, tot = 0;
    for (int k = 0; k < 32; k++) { const u64 a = sm_s[k]; if (k < w) wx += a; tot += a; }
    const u64 kk = sm_carry_k + wx + inc - len;
    float x = FLT_MAX;
    if (on) {
      const float p = __uint_as_float(dk[i]);
      x = __fadd_rn(__fadd_rn(p, logN), log10f_glibc(__ull2float_rn(kk)));
    }
    // inclusive running minimum (over this and larger keys)
    float mn = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(GR_FULL, mn, o);
      if (lane >= o) mn = t < mn ? t : mn;
    }
    if (lane == 31) sm_m[w] = mn;
    __syncthreads();
    float pm = sm_carry_q, all = sm_carry_q;
    for (int k = 0; k < 32; k++) { const float a = sm_m[k]; if (k < w) pm = a < pm ? a : pm; all = a < all ? a : all; }
    mn = pm < mn ? pm : mn;
    if (on) dq[i] = mn > 0.0f ? mn : 0.0f;
    __syncthreads();
    if (threadIdx.x == 0) { sm_carry_k += tot; sm_carry_q = all; }
    __syncthreads();
  }
}

void launch_bh(cudaStream_t s, const u32* keys, const u64* lens, u64 n, float logN, const BhWork& w) {
  cudaMemsetAsync(w.dcount, 0, sizeof(u64), s);
  if (!n) return;
  const u32 nblk = (u32)((n + RS_CHUNK - 1) / RS_CHUNK);
  const u32* kin = keys;
  const u64* lin = lens;
  u32* kbuf[2] = { w.k0, w.k1 };
  u64* lbuf[2] = { w.l0, w.l1 };
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 8 * pass;
    k_rs_hist<<<nblk, RS_THREADS, 0, s>>>(kin, n, shift, w.hist, nblk); GR_NOTE_LAUNCH();
    k_rs_scan<<<1, 1024, 0, s>>>(w.hist, (u64)256 * nblk); GR_NOTE_LAUNCH();
    k_rs_scatter<<<nblk, RS_THREADS, 0, s>>>(kin, lin, n, shift, w.hist, nblk, kbuf[pass & 1], lbuf[pass & 1]); GR_NOTE_LAUNCH();
    kin = kbuf[pass & 1];
    lin = lbuf[pass & 1];
  }
  // sorted data is in k1/l1 (pass 3 -> buffer 1)
  const u32 ntiles = (u32)((n + 255) / 256);
  cudaMemsetAsync(w.sc.st, 0, (size_t)ntiles * sizeof(u64), s);
  cudaMemsetAsync(w.sc.ticket, 0, sizeof(u32), s);
  cudaMemsetAsync(w.dl, 0, n * sizeof(u64), s);
  Lookback<1> lb;
  lb.st[0] = w.sc.st; lb.ticket = w.sc.ticket;
  k_bh_distinct<<<ntiles, 256, 0, s>>>(kin, lin, n, lb, w.dk, w.dl, w.dcount, ntiles); GR_NOTE_LAUNCH();
  k_bh_q<<<1, 1024, 0, s>>>(w.dk, w.dl, w.dcount, logN, w.dq); GR_NOTE_LAUNCH();
}
