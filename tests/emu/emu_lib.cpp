// emu_lib.cpp -- TEST INFRASTRUCTURE: the whole of libgenrich_cuda (kernels AND the host logic of
// gr_api.cu, i.e. every buffer size, launch argument and stage order behind the C-ABI) compiled for
// the host against the lock-step emulation, as one translation unit -> tests/emu/_build/libgenrich_emu.so.
// The CPU test-suite drives it through the same ctypes binding as the CUDA library and compares it
// with the oracle: what a GPU would compute, minus the hardware.  Never loaded by the product.
#include "cuda_runtime.h"      // tests/emu/fake
#include "r_gr_dense.cpp"
#include "r_gr_interval.cpp"
#include "r_gr_bh.cpp"
#include "r_gr_peaks.cpp"
#include "r_gr_api.cpp"
