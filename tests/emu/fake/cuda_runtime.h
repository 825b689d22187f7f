// fake/cuda_runtime.h -- TEST INFRASTRUCTURE: what the product's sources get when they include
// <cuda_runtime.h> in the CPU emulation build (tests/emu): the lock-step execution model
// (cuda_emu.h) and a synchronous stand-in for the part of the CUDA runtime API the library uses.
// "Device" memory is host memory; streams and events do nothing (every launch and copy completes
// before it returns), so stream-ordering mistakes are invisible here -- data flow, indexing, buffer
// sizes and the kernels' arithmetic are not.
#pragma once
#include "cuda_emu.h"
#include <chrono>
#include <map>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct emu_stream_* cudaStream_t;
struct emu_event_ { double t; };
typedef emu_event_* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
static std::map<char*, std::pair<size_t, int>> regions;          // start -> (bytes, type)
static char dyn_smem_buf[256 * 1024] __attribute__((aligned(128)));
static char* dyn_smem = dyn_smem_buf;
static inline int sm_count() { const char* e = getenv("EMU_SMS"); return e ? atoi(e) : 2; }
// EMU_TRACE=1: the name of every kernel as it is launched (which one deadlocked / aborted?)
static inline void trace(const char* kernel) { static const bool on = getenv("EMU_TRACE") != nullptr; if (on) fprintf(stderr, "emu: launch %s\n", kernel); }
template <typename F> static void launch_k(dim3 grid, dim3 block, size_t smem, F body) {
  if (smem > sizeof dyn_smem_buf || grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1) {
    fprintf(stderr, "emu: unsupported launch configuration\n"); abort();
  }
  if (!grid.x || !block.x) return;
  launch(grid.x, block.x, body);
}
}  // namespace emu

// EMU_DEVICES "devices" (default 1) share the one address space, but every device allocation remembers
// the device that was current when it was made, and copies / memsets that touch it from another
// current device abort: a missing cudaSetDevice in a multi-context caller shows up here
static inline int emu_ndev() { const char* e = getenv("EMU_DEVICES"); return e ? atoi(e) : 1; }
static int emu_cur_dev = 0;
namespace emu { static std::map<char*, int> region_dev; }
static inline void emu_check_dev(const void* p, const char* what) {
  auto it = emu::regions.upper_bound((char*)p);
  if (it == emu::regions.begin()) return;
  --it;
  if ((char*)p >= it->first + it->second.first || it->second.second != 2 /* device */) return;
  if (emu::region_dev[it->first] != emu_cur_dev) {
    fprintf(stderr, "emu: %s touches memory of device %d while device %d is current\n", what, emu::region_dev[it->first], emu_cur_dev);
    abort();
  }
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  if (!*p) return cudaErrorMemoryAllocation;
  // device memory is not zeroed: fill it with a pattern (EMU_FILL).  EMU_ZERO_FROM / EMU_ZERO_TO zero the
  // allocations with these ordinal numbers instead -- bisects "which buffer is read before it is written"
  static int ordinal = 0;
  const int k = ordinal++;
  const char* zf = getenv("EMU_ZERO_FROM"); const char* zt = getenv("EMU_ZERO_TO");
  const bool zero = zf && zt && k >= atoi(zf) && k <= atoi(zt);
  memset(*p, zero ? 0 : getenv("EMU_FILL") ? atoi(getenv("EMU_FILL")) : 0xCD, n);
  if (getenv("EMU_ALLOC_LOG")) fprintf(stderr, "emu: cudaMalloc #%d %zu bytes\n", k, n);
  emu::regions[(char*)*p] = {n, cudaMemoryTypeDevice};
  emu::region_dev[(char*)*p] = emu_cur_dev;
  return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { if (p) { emu::regions.erase((char*)p); free(p); } return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  if (!*p) return cudaErrorMemoryAllocation;
  emu::regions[(char*)*p] = {n, cudaMemoryTypeHost};
  return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = a->hostPointer = nullptr;
  auto it = emu::regions.upper_bound((char*)p);
  if (it != emu::regions.begin()) {
    --it;
    if ((char*)p < it->first + it->second.first) a->type = (cudaMemoryType)it->second.second;
  }
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { emu_check_dev(d, "cudaMemcpy"); emu_check_dev(s, "cudaMemcpy"); memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { emu_check_dev(d, "cudaMemcpyAsync"); emu_check_dev(s, "cudaMemcpyAsync"); memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { emu_check_dev(d, "cudaMemset"); memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { emu_check_dev(d, "cudaMemsetAsync"); memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline double emu_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event_{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
static inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu_ndev()) return cudaErrorInvalidValue; emu_cur_dev = d; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = emu_cur_dev; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu_ndev(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = attr == cudaDevAttrMultiProcessorCount ? emu::sm_count() : 0;
  return cudaSuccess;
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
