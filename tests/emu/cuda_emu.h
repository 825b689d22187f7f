// cuda_emu.h -- TEST INFRASTRUCTURE: a lock-step CPU emulation of the CUDA execution model,
// just large enough to run this repository's kernels (as extracted from the .cu files by
// extract.py) on the host cores, so that a kernel written where no GPU is at hand can be checked
// against the kernels already validated on the device.  Nothing in the product includes it.
//
// Model: one CTA at a time; every thread of the CTA is a fiber (own stack, hand-written switch) on ONE OS thread.  A
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
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }
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
struct Group {                             // one rendezvous point: the lanes named by `mask`
  unsigned mask = 0;
  int arrived = 0, pending = 0;            // pending: participants that have not read the last result yet
  unsigned long long gen = 0;
  unsigned long long slot[2][32];          // operands of the collective, double-buffered by generation
  unsigned amask[2] = {0, 0};              // lanes that delivered an operand in that generation
};
struct Fiber {
  void* sp = nullptr;        // saved stack pointer while the fiber is not running
  char* stack = nullptr;
  bool done = false;
  int wait = 0;              // 0 runnable, 1 at a warp rendezvous, 2 at the CTA barrier
  Group* grp = nullptr;
  unsigned long long wgen = 0, bgen = 0;   // generation the fiber waits to see completed
};
struct Warp {
  Group full;                              // all live lanes (the usual case)
  Group part[33];                          // partial masks (__match_any_sync peers, ...), found by mask
  int live = 32;                           // lanes that have not returned
  int init_live = 32;
  unsigned live_mask = 0xffffffffu;
};
struct Cta {
  std::vector<Fiber> f;
  std::vector<Warp> w;
  int nt = 0, live = 0;
  int b_arrived = 0;
  unsigned long long b_gen = 0;
  int cur = -1;
  void* sched_sp = nullptr;
  std::function<void()> body;
};
static Cta* g = nullptr;
static unsigned long long n_switches = 0;
// Context switch between fibers (x86-64 SysV): callee-saved registers and the stack pointer; no signal
// mask round trip (swapcontext makes two system calls per switch, which was 40 % of the run time).
__attribute__((naked, noinline)) static void emu_switch(void** /*save_sp*/, void* /*load_sp*/) {
  asm volatile(
      "pushq %rbp\n\t" "pushq %rbx\n\t" "pushq %r12\n\t" "pushq %r13\n\t" "pushq %r14\n\t" "pushq %r15\n\t"
      "movq %rsp, (%rdi)\n\t"
      "movq %rsi, %rsp\n\t"
      "popq %r15\n\t" "popq %r14\n\t" "popq %r13\n\t" "popq %r12\n\t" "popq %rbx\n\t" "popq %rbp\n\t"
      "ret\n\t");
}
// EMU_ORDER=0 lanes resumed in order, 1 in reverse, 2 in an order that changes from pass to pass: code
// that is only correct because of the order in which the emulation happens to run the lanes
// (a missing __syncwarp / __syncthreads) has three chances to show
static int order_mode = getenv("EMU_ORDER") ? atoi(getenv("EMU_ORDER")) : 0;
static unsigned order_salt = 0;

static inline void release_if_complete(Warp& w, Group& gr) {
  if (gr.arrived > 0 && gr.arrived >= __builtin_popcount(gr.mask & w.live_mask)) {
    gr.pending += gr.arrived; gr.arrived = 0; gr.gen++; gr.amask[gr.gen & 1] = 0;
  }
}
static void entry() {
  g->body();
  Fiber& me = g->f[g->cur];
  me.done = true;
  g->live--;
  Warp& w = g->w[g->cur >> 5];
  w.live--;
  w.live_mask &= ~(1u << (g->cur & 31));
  if (g->live > 0 && g->b_arrived == g->live) { g->b_arrived = 0; g->b_gen++; }   // the others were waiting for this one
  // lanes that have returned do not take part (CUDA: "all non-exited threads named in mask")
  w.full.mask = w.live_mask;
  release_if_complete(w, w.full);
  for (Group& gr : w.part) release_if_complete(w, gr);
  emu_switch(&me.sp, g->sched_sp);
  abort();                                 // a finished fiber is never resumed
}
static bool g_spun = false;                // the fiber that just came back was only polling (spin_yield)
static unsigned long long n_spins = 0;
static inline void yield_();
// a polling loop's yield (an mbarrier wait, a flag): the scheduler moves on to the other warps instead of
// resuming this one at once, and a program in which everybody only polls is reported, not looped forever
static inline void spin_yield() {
  if (++n_spins > 200000000ull) { fprintf(stderr, "emu: every fiber polls -- deadlock\n"); abort(); }
  g_spun = true;
  yield_();
}
static inline void yield_() {
  Fiber& me = g->f[g->cur];
  n_switches++;
  emu_switch(&me.sp, g->sched_sp);
}
// warp rendezvous of the lanes in `mask`; returns the generation (its parity selects the operand
// buffer) and the group, which the caller reads its result from and then leaves
static inline unsigned long long warp_arrive(unsigned long long operand, Group*& gp, unsigned mask = 0xffffffffu) {
  const int t = g->cur;
  Warp& w = g->w[t >> 5];
  mask &= w.live_mask;
  Group* gr = nullptr;
  // a rendezvous of this very mask that is under way comes first: lanes outside the mask may have returned
  // since its first member arrived, and the mask now EQUALS the live mask without being another rendezvous
  for (Group& c : w.part) if (c.mask == mask && (c.arrived || c.pending)) { gr = &c; break; }
  if (gr) {}
  else if (mask == w.live_mask && !(w.full.arrived && w.full.mask != mask)) { gr = &w.full; gr->mask = mask; }
  else {
    if (!gr) for (Group& c : w.part) if (c.mask == mask) { gr = &c; break; }
    if (!gr) for (Group& c : w.part) if (!c.arrived && !c.pending) { gr = &c; c.mask = mask; c.gen = 0; c.amask[0] = c.amask[1] = 0; break; }
    if (!gr) { fprintf(stderr, "emu: out of rendezvous groups\n"); abort(); }
  }
  gp = gr;
  const unsigned long long mygen = gr->gen;
  gr->slot[mygen & 1][t & 31] = operand;
  gr->amask[mygen & 1] |= 1u << (t & 31);
  gr->arrived++;
  if (gr->arrived >= __builtin_popcount(gr->mask & w.live_mask)) {
    gr->pending += gr->arrived; gr->arrived = 0; gr->gen++; gr->amask[gr->gen & 1] = 0;
    return mygen;
  }
  Fiber& me = g->f[t];
  me.wait = 1; me.grp = gr; me.wgen = mygen;
  while (gr->gen == mygen) yield_();
  me.wait = 0;
  return mygen;
}
static inline void warp_leave(Group* gr) { gr->pending--; }
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
  static std::vector<char*> stack_pool;    // stacks are kept across launches (fresh ones cost a page fault per fiber)
  while (stack_pool.size() < nt) stack_pool.push_back((char*)malloc(STACK));
  for (unsigned i = 0; i < nt; i++) cta.f[i].stack = stack_pool[i];
  cta.body = body;
  gridDim = dim3(grid); blockDim = dim3(nt);
  for (unsigned b = 0; b < grid; b++) {
    cta.live = (int)nt; cta.b_arrived = 0;
    for (unsigned i = 0; i < nt; i++) {
      Fiber& f = cta.f[i];
      f.done = false; f.wait = 0;
      // initial frame: six callee-saved registers, then `entry` as the return address of emu_switch;
      // after that `ret` the stack pointer is 8 mod 16, as at any function entry
      void** top = (void**)(((uintptr_t)f.stack + STACK) & ~(uintptr_t)15);
      top[-1] = nullptr;
      top[-2] = (void*)entry;
      for (int k = 3; k <= 8; k++) top[-k] = nullptr;
      f.sp = (void*)(top - 8);
    }
    for (size_t wi = 0; wi < cta.w.size(); wi++) {
      Warp& w = cta.w[wi];
      w.live = w.init_live = (int)std::min<unsigned>(32, nt - (unsigned)wi * 32);
      w.live_mask = w.live == 32 ? 0xffffffffu : ((1u << w.live) - 1u);
      w.full = Group(); w.full.mask = w.live_mask;
      for (Group& gr : w.part) gr = Group();
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
            if (f.wait == 1 && f.grp->gen == f.wgen) continue;        // waits for its siblings
            cta.cur = t;
            threadIdx = uint3{(unsigned)t, 0, 0};
            g_spun = false;
            emu_switch(&cta.sched_sp, f.sp);
            progressed = true;
            if (!g_spun) { any = true; n_spins = 0; }    // a lane that only polled does not keep the scheduler on this warp
          }
          if (order_mode == 2) order_salt = order_salt * 1103515245u + 12345u;
          if (!any) break;
        }
      }
      if (!progressed) {
        fprintf(stderr, "emu: deadlock in CTA %u\n", b);
        if (getenv("EMU_TRACE"))
          for (unsigned t = 0; t < nt; t++) {
            const Fiber& f = cta.f[t];
            if (!f.done) fprintf(stderr, "  thread %u: wait %d%s mask %08x arrived %d pending %d gen %llu (mine %llu)\n", t, f.wait,
                                 f.wait == 1 && f.grp == &cta.w[t >> 5].full ? " (full)" : "", f.wait == 1 ? f.grp->mask : 0u,
                                 f.wait == 1 ? f.grp->arrived : 0, f.wait == 1 ? f.grp->pending : 0,
                                 f.wait == 1 ? f.grp->gen : 0ull, f.wgen);
          }
        abort();
      }
    }
  }
  g = nullptr;
}
}  // namespace emu

// ---- collectives ------------------------------------------------------------------------------
#define emu_lane_in(gr, l) (((gr)->amask[gen & 1] >> (l)) & 1u)      /* lane l delivered an operand to this rendezvous */
static inline void __syncthreads() { emu::cta_barrier(); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { emu::Group* w; emu::warp_arrive(0, w, m); emu::warp_leave(w); }
template <typename T> static inline T emu_bits_to(unsigned long long v) { T r; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline unsigned long long emu_to_bits(T v) { unsigned long long r = 0; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) {
  emu::Group* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w, m);
  const T r = emu_bits_to<T>(w->slot[gen & 1][src & 31]);
  emu::warp_leave(w);
  return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Group* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w, m);
  const T r = lane >= (int)d ? emu_bits_to<T>(w->slot[gen & 1][lane - (int)d]) : v;
  emu::warp_leave(w);
  return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Group* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w, m);
  const T r = lane + (int)d < 32 ? emu_bits_to<T>(w->slot[gen & 1][lane + (int)d]) : v;
  emu::warp_leave(w);
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) {
  const int lane = (int)(threadIdx.x & 31);
  emu::Group* w; const unsigned long long gen = emu::warp_arrive(emu_to_bits(v), w, m);
  const T r = emu_bits_to<T>(w->slot[gen & 1][(lane ^ x) & 31]);
  emu::warp_leave(w);
  return r;
}
#define EMU_REDUCE(init, expr)                                                          \
  emu::Group* w; const unsigned long long gen = emu::warp_arrive(operand, w, m);          \
  unsigned r = init;                                                                    \
  for (int l = 0; l < 32; l++) if (emu_lane_in(w, l)) { const unsigned long long x = w->slot[gen & 1][l]; (void)x; expr; } \
  emu::warp_leave(w);                                                                   \
  return r;
static inline unsigned __ballot_sync(unsigned m, bool p) { const unsigned long long operand = p ? 1ull : 0ull; EMU_REDUCE(0u, r |= (unsigned)(x & 1ull) << l) }
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0; }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { const unsigned long long operand = v; EMU_REDUCE(0u, r += (unsigned)x) }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { const unsigned long long operand = v; EMU_REDUCE(0u, r = std::max(r, (unsigned)x)) }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { const unsigned long long operand = v; EMU_REDUCE(0u, r |= (unsigned)x) }
static inline unsigned __match_any_sync(unsigned m, unsigned long long v) { const unsigned long long operand = v; EMU_REDUCE(0u, r |= (unsigned)(x == operand) << l) }
