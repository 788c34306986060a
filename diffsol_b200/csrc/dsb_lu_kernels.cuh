// dsb_lu_kernels.cuh -- the `LinearSolver<M>` pair as stand-alone batched kernels.
//
// Replaces `NalgebraLU::set_linearisation` / `solve_in_place`
// (crates/diffsol-la/src/linear_solver/nalgebra/lu.rs:31-51) and the reference's per-instance
// cuSOLVER loop (crates/diffsol-la/src/linear_solver/cuda/lu.rs:80-95,127-145) with one launch for the
// whole batch.  Same arithmetic as nalgebra's partial-pivot LU (first maximum, reciprocal-pivot
// scaling, column-axpy updates and substitutions), so factors and solutions are bit-identical to the
// CPU path.  Layout is batch-major: a[(j*n + i)*B + b], piv[i*B + b], rhs[i*B + b]; one thread per
// instance, so every load/store of a warp is one coalesced 256-byte segment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsb_coop.cuh"
#include "dsb_lane.cuh"

// Small systems: the whole matrix in registers (static indices after unrolling).
template <int N>
__global__ void __launch_bounds__(128) dsb_lu_factor_reg_kernel(double* __restrict__ a, int64_t B, int32_t* __restrict__ piv,
                                                                int32_t* __restrict__ info) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    LaneLU<N> lu;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) lu.a[j][i] = a[((int64_t)j * N + i) * B + b];
    lu.factor();
    int bad = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[((int64_t)j * N + i) * B + b] = lu.a[j][i];
        piv[(int64_t)j * B + b] = lu.piv[j];
        if (lu.a[j][j] == 0.0 && bad == 0) bad = j + 1;
    }
    info[b] = bad;
}
template <int N>
__global__ void __launch_bounds__(128) dsb_lu_solve_reg_kernel(const double* __restrict__ a, const int32_t* __restrict__ piv,
                                                               double* __restrict__ rhs, int64_t B, int32_t* __restrict__ info) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    LaneLU<N> lu;
    double x[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) lu.a[j][i] = a[((int64_t)j * N + i) * B + b];
        lu.piv[j] = piv[(int64_t)j * B + b];
        x[j] = rhs[(int64_t)j * B + b];
    }
    const bool ok = lu.solve(x);
#pragma unroll
    for (int j = 0; j < N; ++j) rhs[(int64_t)j * B + b] = x[j];
    info[b] = ok ? 0 : 1;
}

// General n: the matrix stays in global memory (L1/L2 keep the instance's working set hot; accesses
// remain coalesced across the warp because every lane walks the same (i, j) sequence).
__global__ void __launch_bounds__(128) dsb_lu_factor_gmem_kernel(double* __restrict__ a, int n, int64_t B,
                                                                 int32_t* __restrict__ piv, int32_t* __restrict__ info) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    auto at = [&](int i, int j) -> double& { return a[((int64_t)j * n + i) * B + b]; };
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        int p = i;
        double the_max = dsb_abs(at(i, i));
        for (int r = i + 1; r < n; ++r) {
            const double val = dsb_abs(at(r, i));
            if (val > the_max) { the_max = val; p = r; }
        }
        const double diag = at(p, i);
        if (diag == 0.0) { piv[(int64_t)i * B + b] = i; if (!bad) bad = i + 1; continue; }
        piv[(int64_t)i * B + b] = p;
        if (p != i) {
            for (int c = 0; c < n; ++c) { const double tmp = at(i, c); at(i, c) = at(p, c); at(p, c) = tmp; }
        }
        const double inv_diag = 1.0 / diag;
        for (int r = i + 1; r < n; ++r) at(r, i) *= inv_diag;
        for (int k = i + 1; k < n; ++k) {
            const double mpk = -at(i, k);
            for (int r = i + 1; r < n; ++r) at(r, k) = mpk * at(r, i) + at(r, k);
        }
    }
    info[b] = bad;
}
__global__ void __launch_bounds__(128) dsb_lu_solve_gmem_kernel(const double* __restrict__ a, const int32_t* __restrict__ piv,
                                                                double* __restrict__ rhs, int n, int64_t B,
                                                                int32_t* __restrict__ info) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    auto at = [&](int i, int j) -> double { return a[((int64_t)j * n + i) * B + b]; };
    auto x = [&](int i) -> double& { return rhs[(int64_t)i * B + b]; };
    for (int i = 0; i < n; ++i) {
        const int p = piv[(int64_t)i * B + b];
        if (p != i) { const double tmp = x(i); x(i) = x(p); x(p) = tmp; }
    }
    for (int i = 0; i + 1 < n; ++i) {
        const double mc = -(x(i) / 1.0);
        for (int r = i + 1; r < n; ++r) x(r) = mc * at(r, i) + x(r);
    }
    int bad = 0;
    for (int i = n - 1; i >= 0; --i) {
        const double diag = at(i, i);
        if (diag == 0.0) { bad = 1; break; }
        const double coeff = x(i) / diag;
        x(i) = coeff;
        const double mc = -coeff;
        for (int r = 0; r < i; ++r) x(r) = mc * at(r, i) + x(r);
    }
    info[b] = bad;
}

inline cudaError_t dsb_launch_lu_factor(double* a, int n, int64_t B, int32_t* piv, int32_t* info, cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((B + threads - 1) / threads);
    switch (n) {
        case 1: dsb_lu_factor_reg_kernel<1><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 2: dsb_lu_factor_reg_kernel<2><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 3: dsb_lu_factor_reg_kernel<3><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 4: dsb_lu_factor_reg_kernel<4><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 5: dsb_lu_factor_reg_kernel<5><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 6: dsb_lu_factor_reg_kernel<6><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 7: dsb_lu_factor_reg_kernel<7><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        case 8: dsb_lu_factor_reg_kernel<8><<<blocks, threads, 0, s>>>(a, B, piv, info); break;
        default: dsb_lu_factor_gmem_kernel<<<blocks, threads, 0, s>>>(a, n, B, piv, info); break;
    }
    return cudaGetLastError();
}
inline cudaError_t dsb_launch_lu_solve(const double* a, const int32_t* piv, double* rhs, int n, int64_t B, int32_t* info,
                                       cudaStream_t s) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((B + threads - 1) / threads);
    switch (n) {
        case 1: dsb_lu_solve_reg_kernel<1><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 2: dsb_lu_solve_reg_kernel<2><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 3: dsb_lu_solve_reg_kernel<3><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 4: dsb_lu_solve_reg_kernel<4><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 5: dsb_lu_solve_reg_kernel<5><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 6: dsb_lu_solve_reg_kernel<6><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 7: dsb_lu_solve_reg_kernel<7><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        case 8: dsb_lu_solve_reg_kernel<8><<<blocks, threads, 0, s>>>(a, piv, rhs, B, info); break;
        default: dsb_lu_solve_gmem_kernel<<<blocks, threads, 0, s>>>(a, piv, rhs, n, B, info); break;
    }
    return cudaGetLastError();
}

// ---- stand-alone kernels: the LinearSolver pair for instance-major storage ------------------------------
// a: [nbatch][n*n] column-major per instance (the layout of the reference's CUDA matrices,
// diffsol-la/src/matrix/cuda.rs), piv: [nbatch][n], rhs: [nbatch][n].
__global__ void __launch_bounds__(128, 3) dsb_lu_factor_coop_kernel(double* __restrict__ a, int n, int64_t B, int32_t* __restrict__ piv,
                                          int32_t* __restrict__ info) {
    extern __shared__ unsigned char dsb_coop_smem[];
    const CoopScratch sc = coop_carve(dsb_coop_smem, n);
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const int bad = coop_lu_factor(a + (size_t)b * n * n, n, piv + (size_t)b * n, sc);
        if (threadIdx.x == 0) info[b] = bad;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(128, 4) dsb_lu_solve_coop_kernel(const double* __restrict__ a, const int32_t* __restrict__ piv,
                                         double* __restrict__ rhs, int n, int64_t B, int32_t* __restrict__ info) {
    extern __shared__ unsigned char dsb_coop_smem[];
    CoopScratch sc;                                      // no panel here: [b (n doubles) | reduction scratch]
    sc.panel = (double*)dsb_coop_smem;
    sc.utile = nullptr;
    sc.redv = sc.panel + n; sc.redi = (int*)(sc.redv + 32); sc.bcast = sc.redi + 32;
    double* bs = sc.panel;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) bs[i] = rhs[(size_t)b * n + i];
        __syncthreads();
        const bool ok = coop_lu_solve(a + (size_t)b * n * n, n, piv + (size_t)b * n, bs, sc, false);     // no panel scratch in this kernel
        for (int i = threadIdx.x; i < n; i += blockDim.x) rhs[(size_t)b * n + i] = bs[i];
        if (threadIdx.x == 0) info[b] = ok ? 0 : 1;
        __syncthreads();
    }
}

// ---- launchers for the instance-major kernels ----
inline int dsb_coop_threads(int n) { int t = ((n + 31) / 32) * 32; return t > 128 ? 128 : t; }
inline cudaError_t dsb_coop_grid(const void* kernel, int threads, size_t smem, int64_t B, unsigned* grid) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int64_t resident = (int64_t)sms * per_sm;
    *grid = (unsigned)(B < resident ? B : resident);
    return cudaSuccess;
}
inline cudaError_t dsb_launch_lu_factor_im(double* a, int n, int64_t B, int32_t* piv, int32_t* info, cudaStream_t s) {
    if (n > DSB_COOP_MAX_N) return cudaErrorInvalidValue;
    const int threads = dsb_coop_threads(n);
    const size_t smem = coop_lu_smem_bytes_host(n);
    unsigned grid = 1;
    cudaError_t e = dsb_coop_grid((const void*)dsb_lu_factor_coop_kernel, threads, smem, B, &grid);
    if (e != cudaSuccess) return e;
    dsb_lu_factor_coop_kernel<<<grid, threads, smem, s>>>(a, n, B, piv, info);
    return cudaGetLastError();
}
inline cudaError_t dsb_launch_lu_solve_im(const double* a, const int32_t* piv, double* rhs, int n, int64_t B, int32_t* info,
                                          cudaStream_t s) {
    if (n > DSB_COOP_MAX_N) return cudaErrorInvalidValue;
    const int threads = dsb_coop_threads(n);
    const size_t smem = coop_lu_smem_bytes_host(n) - ((size_t)n * (DSB_COOP_NB - 1) + DSB_COOP_UT_WORDS) * sizeof(double);   // only b: no panel, no U tiles
    unsigned grid = 1;
    cudaError_t e = dsb_coop_grid((const void*)dsb_lu_solve_coop_kernel, threads, smem, B, &grid);
    if (e != cudaSuccess) return e;
    dsb_lu_solve_coop_kernel<<<grid, threads, smem, s>>>(a, piv, rhs, n, B, info);
    return cudaGetLastError();
}
