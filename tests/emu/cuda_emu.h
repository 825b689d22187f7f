// cuda_emu.h -- TEST INFRASTRUCTURE: a lock-step CPU emulation of the CUDA execution model,
// just large enough to run this repository's kernels (as extracted from the .cu files by
// extract.py) on the host cores, so that a kernel written where no GPU is at hand can be checked
// against the kernels already validated on the device.  Nothing in the product includes it.
//
// Model: one CTA at a time; every thread of the CTA is a ucontext fiber on ONE OS thread.  A
// warp-level collective (__shfl*_sync, __ballot_sync, __reduce_*_sync, __syncwarp) or a CTA
// barrier (__syncthreads) is a rendezvous: the fiber publishes its operand, yields, and is
// resumed once every participant has arrived.  The scheduler runs the lanes of one warp
// round-robin until all of them wait at a CTA barrier (or have returned), then the next warp.
// Full masks only (all this code base uses).  Shared memory: `__shared__` becomes `static`
// (CTAs run one after the other, so one copy per kernel instantiation is enough); atomics are
// plain read-modify-writes.  Floating point: compile with -ffp-contract=off.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <algorithm>
#include <functional>
#include <vector>

#define GR_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __constant__ static

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

using std::min;
using std::max;
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// ---- per-thread built-ins (set by the scheduler at every switch) --------------------------
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

// ---- bit / conversion intrinsics ---------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) {}

// ---- atomics: one OS thread, so plain read-modify-writes ---------------------------------
template <typename T, typename V> static inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename V> static inline T atomicOr(T* p, V v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename V> static inline T atomicAnd(T* p, V v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename V> static inline T atomicMax(T* p, V v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename V> static inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename V> static inline T atomicExch(T* p, V v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename C, typename V> static inline T atomicCAS(T* p, C c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// ---- the scheduler -------------------------------------------------------------------------
namespace emu {
enum { STACK = 256 * 1024 };
struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  int wait = 0;              // 0 runnable, 1 at a warp rendezvous, 2 at the CTA barrier
  unsigned long long wgen = 0, bgen = 0;   // generation the fiber waits to see completed
};
struct Warp {
  int arrived = 0;
  unsigned long long gen = 0;
  unsigned long long slot[2][32];          // operands of the collective, double-buffered by generation
  int live = 32;                           // lanes that have not returned
  int init_live = 32;
};
struct Cta {
  std::vector<Fiber> f;
  std::vector<Warp> w;
  int nt = 0, live = 0;
  int b_arrived = 0;
  unsigned long long b_gen = 0;
  int cur = -1;
  ucontext_t sched;
  std::function<void()> body;
};
static Cta* g = nullptr;
static unsigned long long n_switches = 0;
// EMU_ORDER=0 lanes resumed in order, 1 in reverse, 2 in an order that changes from pass to pass: code
// that is only correct because of the order in which the emulation happens to run the lanes
// (a missing __syncwarp / __syncthreads) has three chances to show
static int order_mode = getenv("EMU_ORDER") ? atoi(getenv("EMU_ORDER")) : 0;
static unsigned order_salt = 0;

static void entry() {
  g->body();
  Fiber& me = g->f[g->cur];
  me.done = true;
  g->live--;
  Warp& w = g->w[g->cur >> 5];
  w.live--;
  if (g->live > 0 && g->b_arrived == g->live) { g->b_arrived = 0; g->b_gen++; }   // the others were waiting for this one
  // a lane that returns while its siblings wait must not leave them hanging: the code base only
  // returns warp-uniformly before collectives, which the counters below would flag otherwise
  swapcontext(&me.ctx, &g->sched);
}
static inline void yield_() {
  Fiber& me = g->f[g->cur];
  n_switches++;
  swapcontext(&me.ctx, &g->sched);
}
// warp rendezvous; returns the generation index (parity selects the operand buffer)
static inline unsigned long long warp_arrive(unsigned long long operand, Warp*& wp) {
  const int t = g->cur;
  Warp& w = g->w[t >> 5];
  wp = &w;
  const unsigned long long mygen = w.gen;
  w.slot[mygen & 1][t & 31] = operand;
  if (w.live != w.init_live) {
    fprintf(stderr, "emu: warp collective after a divergent return (thread %d)\n", t); abort();
  }
  w.arrived++;
  if (w.arrived == w.live) { w.arrived = 0; w.gen++; return mygen; }
  Fiber& me = g->f[t];
  me.wait = 1; me.wgen = mygen;
  while (w.gen == mygen) yield_();
  me.wait = 0;
  return mygen;
}
static inline void cta_barrier() {
  const int t = g->cur;
  const unsigned long long mygen = g->b_gen;
  g->b_arrived++;
  if (g->b_arrived == g->live) { g->b_arrived = 0; g->b_gen++; return; }
  Fiber& me = g->f[t];
  me.wait = 2; me.bgen = mygen;
  while (g->b_gen == mygen) yield_();
  me.wait = 0;
}

// run `body` as a grid of CTAs of nt threads
template <typename F> static void launch(unsigned grid, unsigned nt, F body) {
  Cta cta;
  g = &cta;
  cta.nt = (int)nt;
  cta.f.resize(nt);
  cta.w.resize((nt + 31) / 32);
  for (auto& f : cta.f) f.stack = (char*)malloc(STACK);
  cta.body = body;
  gridDim = dim3(grid); blockDim = dim3(nt);
  for (unsigned b = 0; b < grid; b++) {
    cta.live = (int)nt; cta.b_arrived = 0;
    for (unsigned i = 0; i < nt; i++) {
      Fiber& f = cta.f[i];
      f.done = false; f.wait = 0;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = STACK;
      f.ctx.uc_link = &cta.sched;
      makecontext(&f.ctx, (void (*)())entry, 0);
    }
    for (size_t wi = 0; wi < cta.w.size(); wi++) {
      cta.w[wi].arrived = 0;
      cta.w[wi].live = cta.w[wi].init_live = (int)std::min<unsigned>(32, nt - (unsigned)wi * 32);
    }
    blockIdx = uint3{b, 0, 0};
    while (cta.live > 0) {
      bool progressed = false;
      for (size_t wi = 0; wi < cta.w.size(); wi++) {
        // the lanes of this warp, round-robin, until every one of them is done or parked at the CTA barrier
        for (;;) {
          bool any = false;
          for (int l0 = 0; l0 < 32; l0++) {
            const int l = order_mode == 0 ? l0 : order_mode == 1 ? 31 - l0 : (int)((l0 * 13 + order_salt) & 31);
            const int t = (int)wi * 32 + l;
            if (t >= (int)nt) continue;
            Fiber& f = cta.f[t];
            if (f.done) continue;
            if (f.wait == 2 && cta.b_gen == f.bgen) continue;         // parked at the CTA barrier
            if (f.wait == 1 && cta.w[wi].gen == f.wgen) continue;     // waits for its siblings
            cta.cur = t;
            threadIdx = uint3{(unsigned)t, 0, 0};
            swapcontext(&cta.sched, &f.ctx);
            any = true; progressed = true;
          }
          if (order_mode == 2) order_salt = order_salt * 1103515245u + 12345u;
          if (!any) break;
        }
      }
      if (!progressed) { fprintf(stderr, "emu: deadlock in CTA %u\n", b); abort(); }
    }
  }
  for (auto& f : cta.f) free(f.stack);
  g = nullptr;
}
}  // namespace emu

// ---- collectives (full mask) ---------------------------------------------------------------
static inline void __syncthreads() { emu::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::Warp* w; emu::warp_arrive(0, w); }
template <typename T> static inline T emu_bits_to(unsigned long long v) { T r; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline unsigned long long emu_to_bits(T v) { unsigned long long r = 0; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w);
  return emu_bits_to<T>(w->slot[gen & 1][src & 31]);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w);
  return lane >= (int)d ? emu_bits_to<T>(w->slot[gen & 1][lane - (int)d]) : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w);
  return lane + (int)d < 32 ? emu_bits_to<T>(w->slot[gen & 1][lane + (int)d]) : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w);
  return emu_bits_to<T>(w->slot[gen & 1][(lane ^ m) & 31]);
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(p ? 1ull : 0ull, w);
  unsigned r = 0;
  for (int l = 0; l < w->live; l++) r |= (unsigned)(w->slot[gen & 1][l] & 1ull) << l;
  return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(v, w);
  unsigned r = 0;
  for (int l = 0; l < w->live; l++) r += (unsigned)w->slot[gen & 1][l];
  return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(v, w);
  unsigned r = 0;
  for (int l = 0; l < w->live; l++) r = std::max(r, (unsigned)w->slot[gen & 1][l]);
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  emu::Warp* w; const unsigned long long gen = emu::warp_arrive(v, w);
  unsigned r = 0;
  for (int l = 0; l < w->live; l++) r |= (unsigned)w->slot[gen & 1][l];
  return r;
}
