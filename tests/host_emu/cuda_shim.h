// cuda_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the one-thread-per-instance ("lane") kernels of
// diffsol_b200/csrc/*.cuh unchanged and run ONE lane of them on the host, so that the kernels' control flow and
// arithmetic can be compared bit for bit with the oracle on machines without a GPU (`pytest -m "not gpu"`).
// What is emulated: a grid of one block of one thread; warp votes answer as if all 32 lanes were in the emulated
// lane's state (the kernels' scheduler only decides WHEN a lane runs a block, never what it computes).  Nothing
// under diffsol_b200/ includes this file; the product has no host integrator.
#pragma once
#include <stdint.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __maxnreg__(x)
#define __launch_bounds__(...)
#define __shared__

struct dsb_emu_dim3 { unsigned x, y, z; };
static dsb_emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 0xffffffffu : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline bool __all_sync(unsigned, bool p) { return p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __syncthreads_and(int p) { return p; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }

// the division policies of dsb_math.h (device-only there)
struct DsbDivShared { static inline double div(double a, double b) { return a / b; } };
struct DsbDivInline { static inline double div(double a, double b) { return a / b; } };
