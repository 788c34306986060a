// dsb_inst.cu -- instantiates the kernel families for ONE equation set.
//   built-in equation sets: compile with -DDSB_INST=<model id> (one translation unit per model, diffsol_b200/build.py);
//   USER equation sets, at run time (dsb_capi.cu: dsb_model_library_build): -DDSB_USER_MODEL_SOURCE="<file>" plus either
//   -DDSB_USER_MODEL=<struct name> (a functor with the interface of dsb_models.h) or -DDSB_USER_DIFFSL (the DiffSL symbol
//   table, dsb_diffsl_adapter.h) -- the same kernels, exported from a plugin shared object as dsb_plugin_* (bottom).
#include <cmath>
#include <cstdlib>
#include <limits>
#include <vector>

#include "dsb_band_bdf_kernel.cuh"
#include "dsb_band_init_kernel.cuh"
#include "dsb_band_sdirk_kernel.cuh"
#include "dsb_bdf_kernel.cuh"
#include "dsb_coop_bdf_kernel.cuh"
#include "dsb_host_setup.h"
#include "dsb_init_kernel.cuh"
#include "dsb_launch.h"
#include "dsb_models.h"
#include "dsb_sdirk_kernel.cuh"
#include "dsb_wband_bdf_kernel.cuh"

#define DSB_CAT_(a, b) a##b
#define DSB_CAT(a, b) DSB_CAT_(a, b)

#if defined(DSB_USER_MODEL_SOURCE)
#include DSB_USER_MODEL_SOURCE
#if defined(DSB_USER_DIFFSL)
#include "dsb_diffsl_adapter.h"
typedef DsbDiffslModel InstModel;
#else
typedef DSB_USER_MODEL InstModel;
#endif
#define DSB_LAUNCH_SYMBOL dsb_plugin_launch
#else
#ifndef DSB_INST
#error "compile with -DDSB_INST=<model id>"
#endif
typedef dsb_model_by_id<DSB_INST>::type InstModel;
#define DSB_LAUNCH_SYMBOL DSB_CAT(dsb_launch_model_, DSB_INST)
#endif

constexpr bool kLaneCapable = InstModel::N <= 16;

template <class M>
static cudaError_t launch_coop_bdf(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, cudaStream_t stream,
                                   cudaEvent_t mid, unsigned long long* work_counter, DsbCoopState* coop,
                                   const double* atol_host, int* launches, bool rk) {
    constexpr int N = M::N;
    // one kernel, two instantiations: Bdf, or Sdirk with the tableau of pa->rk (TR-BDF2 / ESDIRK34)
    void (*const kern)(const DsbProblemArgs, const DsbBatchBuffers, const DsbCoopWorkspace, unsigned long long*) =
        rk ? dsb_coop_bdf_solve_dense_kernel<M, true> : dsb_coop_bdf_solve_dense_kernel<M, false>;
    const int threads = CoopBdfLayout<M>::THREADS;
    const size_t smem = CoopBdfLayout<M>::smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int64_t resident = (int64_t)sms * per_sm;
    const unsigned grid = (unsigned)(pa->nbatch < resident ? pa->nbatch : resident);
    // workspace: per resident block J, [M], LU (n^2 doubles each) and n pivots
    const size_t nn = (size_t)N * N;
    const size_t per_block = (M::HAS_MASS ? 3 : 2) * nn * sizeof(double) + (size_t)N * sizeof(int32_t);
    const size_t need = per_block * grid + 256;
    if (coop->ws_bytes < need) {
        if (coop->ws_mem) cudaFree(coop->ws_mem);
        coop->ws_mem = nullptr; coop->ws_bytes = 0;
        e = cudaMalloc(&coop->ws_mem, need);
        if (e != cudaSuccess) return e;
        coop->ws_bytes = need;
    }
    if (coop->atol_n < N) {
        if (coop->atol_dev) cudaFree(coop->atol_dev);
        coop->atol_dev = nullptr; coop->atol_n = 0;
        e = cudaMalloc((void**)&coop->atol_dev, (size_t)N * sizeof(double));
        if (e != cudaSuccess) return e;
        coop->atol_n = N;
    }
    e = cudaMemcpyAsync(coop->atol_dev, atol_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    DsbCoopWorkspace ws;
    char* base = (char*)coop->ws_mem;
    ws.jac = (double*)base; base += nn * sizeof(double) * grid;
    ws.mass = nullptr;
    if (M::HAS_MASS) { ws.mass = (double*)base; base += nn * sizeof(double) * grid; }
    ws.lu = (double*)base; base += nn * sizeof(double) * grid;
    ws.piv = (int32_t*)base;
    ws.atol = coop->atol_dev;
    ws.color = nullptr; ws.nz = nullptr;
    if (pa->use_coloring) {
        if (!coop->color_host || !coop->nz_host) return cudaErrorInvalidValue;
        const size_t cbytes = (size_t)N * sizeof(int32_t) + nn;
        if (coop->color_bytes < cbytes) {
            if (coop->color_dev) cudaFree(coop->color_dev);
            coop->color_dev = nullptr; coop->color_bytes = 0;
            e = cudaMalloc(&coop->color_dev, cbytes);
            if (e != cudaSuccess) return e;
            coop->color_bytes = cbytes;
        }
        e = cudaMemcpyAsync(coop->color_dev, coop->color_host, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync((char*)coop->color_dev + (size_t)N * sizeof(int32_t), coop->nz_host, nn, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);          // the host vectors die with the caller's frame
        if (e != cudaSuccess) return e;
        ws.color = (const int32_t*)coop->color_dev;
        ws.nz = (const uint8_t*)((char*)coop->color_dev + (size_t)N * sizeof(int32_t));
    }
    e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (mid) cudaEventRecord(mid, stream);
    kern<<<grid, threads, smem, stream>>>(*pa, *bb, ws, work_counter);
    *launches += 1;
    return cudaGetLastError();
}

// A model plugin of a small system (n <= 16) carries the on-chip lane kernels only: the block-per-instance kernel is the
// largest of the families and would double the run-time compilation for a path such a model never takes by default.
#if defined(DSB_USER_MODEL_SOURCE)
constexpr bool kCoopBuilt = !kLaneCapable;
#else
constexpr bool kCoopBuilt = true;
#endif
template <class M, bool BUILT> struct CoopLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, cudaStream_t, cudaEvent_t, unsigned long long*, DsbCoopState*,
                           const double*, int*, bool) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct CoopLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, DsbCoopState* coop, const double* atol_host, int* launches, bool rk) {
        return launch_coop_bdf<M>(pa, bb, stream, mid, work_counter, coop, atol_host, launches, rk);
    }
};

// banded lane kernel (one thread per instance, state in global memory): component-wise models without a mass
// matrix that declare a band, n > 16
constexpr bool kBandCapable = dsb_declares_band<InstModel>::value && dsb_is_componentwise<InstModel>::value && InstModel::N > 16;

template <class M, bool BAND> struct BandLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, int, cudaStream_t, cudaEvent_t, unsigned long long*, DsbCoopState*,
                           const double*, int*) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct BandLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, DsbCoopState* coop, const double* atol_host, int* launches) {
        typedef BandBdfLayout<M, DSB_BAND_THREADS> Lay;
        constexpr int N = M::N;
        // sparsity pattern by NaN probe (jacobian/mod.rs:16-48): the declared band must cover it.  Per column: colour
        // (greedy colouring computed by the caller, or one colour per column for the dense assembly, which stores
        // every in-band entry) and the pattern inside the band.
        std::vector<int32_t> colmeta;
        if (pa->use_coloring && !coop->color_host) return cudaErrorInvalidValue;
        if (!dsb_host::band_column_meta<M>(pa->t0, pa->use_coloring != 0, coop->color_host, Lay::KL, Lay::KU, &colmeta))
            return cudaErrorNotSupported;
        cudaError_t e;
        if (coop->atol_n < N) {
            if (coop->atol_dev) cudaFree(coop->atol_dev);
            coop->atol_dev = nullptr; coop->atol_n = 0;
            e = cudaMalloc((void**)&coop->atol_dev, (size_t)N * sizeof(double));
            if (e != cudaSuccess) return e;
            coop->atol_n = N;
        }
        if (coop->color_bytes < (size_t)N * sizeof(int32_t)) {
            if (coop->color_dev) cudaFree(coop->color_dev);
            coop->color_dev = nullptr; coop->color_bytes = 0;
            e = cudaMalloc(&coop->color_dev, (size_t)N * sizeof(int32_t));
            if (e != cudaSuccess) return e;
            coop->color_bytes = (size_t)N * sizeof(int32_t);
        }
        e = cudaMemcpyAsync(coop->atol_dev, atol_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(coop->color_dev, colmeta.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);              // the host vectors die with this frame
        if (e != cudaSuccess) return e;
        const DsbBandMeta meta{coop->atol_dev, (const int32_t*)coop->color_dev};
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        // a batch that does not fill one big block per SM is spread over the SMs in small blocks
        bool small = (pa->nbatch + DSB_BAND_THREADS - 1) / DSB_BAND_THREADS < sms;
        if (const char* q = getenv("DSB_BAND_BLOCK")) {           // test / tuning hook: force the block shape
            const int v = atoi(q);
            if (v == DSB_BAND_THREADS) small = false; else if (v == DSB_BAND_THREADS_SMALL) small = true;
        }
        if (small)
            return launch<DSB_BAND_THREADS_SMALL>(pa, bb, method, stream, mid, work_counter, coop, meta, sms, launches);
        return launch<DSB_BAND_THREADS>(pa, bb, method, stream, mid, work_counter, coop, meta, sms, launches);
    }
    template <int T>
    static cudaError_t launch(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method, cudaStream_t stream, cudaEvent_t mid,
                              unsigned long long* work_counter, DsbCoopState* coop, const DsbBandMeta& meta, int sms, int* launches) {
        typedef BandBdfLayout<M, T> LayB;
        typedef BandSdirkLayout<M, T> LayS;
        const bool bdf = method == DSB_METHOD_BDF;
        const int threads = T;
        const size_t smem = (size_t)(bdf ? LayB::SMEM_WORDS : LayS::SMEM_WORDS) * threads * sizeof(double);
        const void* kernel = bdf ? (const void*)dsb_band_bdf_solve_dense_kernel<M, T> : (const void*)dsb_band_sdirk_solve_dense_kernel<M, T>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        const int64_t want = (pa->nbatch + threads - 1) / threads;
        const int64_t resident = (int64_t)sms * per_sm;
        const unsigned grid = (unsigned)(want < resident ? want : resident);
        // workspace: one column of WORDS doubles per resident lane (LayS::WORDS covers both kernels and the initialisation)
        const size_t need = (size_t)LayS::WORDS * grid * threads * sizeof(double);
        if (coop->ws_bytes < need) {
            if (coop->ws_mem) cudaFree(coop->ws_mem);
            coop->ws_mem = nullptr; coop->ws_bytes = 0;
            e = cudaMalloc(&coop->ws_mem, need);
            if (e != cudaSuccess) return e;
            coop->ws_bytes = need;
        }
        e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        if constexpr (M::HAS_MASS) {        // consistent initialisation of the DAE on the same lanes and workspace
            dsb_band_init_kernel<M, T><<<grid, threads, 0, stream>>>(*pa, *bb, meta, (double*)coop->ws_mem);
            *launches += 1;
        }
        if (mid) cudaEventRecord(mid, stream);
        if (bdf) dsb_band_bdf_solve_dense_kernel<M, T><<<grid, threads, smem, stream>>>(*pa, *bb, meta, (double*)coop->ws_mem, work_counter);
        else dsb_band_sdirk_solve_dense_kernel<M, T><<<grid, threads, smem, stream>>>(*pa, *bb, meta, (double*)coop->ws_mem, work_counter);
        *launches += 1;
        return cudaGetLastError();
    }
};

// warp-per-instance banded kernel (dsb_wband_bdf_kernel.cuh): BDF, band-capable equations without a reset function
template <class M, bool BAND> struct WBandCapable : std::false_type {};
template <class M> struct WBandCapable<M, true> : std::bool_constant<!dsb_model_has_reset<M>::value && WBandLayout<M>::FITS> {};
constexpr bool kWBandCapable = WBandCapable<InstModel, kBandCapable>::value;
// automatic selection between the two banded kernel families, from measurements on one B200 (250 000 instances, BDF; ms
// warp-per-instance vs lane-per-instance): n = 256 DAE 100 vs 374; n = 200 616 vs 1113, with out / stop functions 1180 vs
// 1556; n = 42 182 vs 231, but with out / stop functions 731 vs 514 -- a warp evaluates its instance's output and root
// functions once per step at the issue cost of 32 lanes, which a small system cannot amortise.  Hence: the warp kernel
// for every larger system (difference array in its global-memory slot) and for small systems without output / root
// functions; one lane per instance for small systems that have them.
template <class M, bool OK> struct WBandPreferred : std::false_type {};
template <class M> struct WBandPreferred<M, true>
    : std::bool_constant<!WBandLayout<M>::D_SHARED || (dsb_model_nroots<M>::value == 0 && !dsb_model_nout<M>::has_out)> {};
constexpr bool kWBandPreferred = WBandPreferred<InstModel, kWBandCapable>::value;

template <class M, bool OK> struct WBandLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, cudaStream_t, cudaEvent_t, unsigned long long*, DsbCoopState*,
                           const double*, int*) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct WBandLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, DsbCoopState* coop, const double* atol_host, int* launches) {
        typedef WBandLayout<M> Lay;
        constexpr int N = M::N;
        std::vector<int32_t> colmeta;
        if (pa->use_coloring && !coop->color_host) return cudaErrorInvalidValue;
        if (!dsb_host::band_column_meta<M>(pa->t0, pa->use_coloring != 0, coop->color_host, Lay::KL, Lay::KU, &colmeta))
            return cudaErrorNotSupported;
        cudaError_t e;
        if (coop->atol_n < N) {
            if (coop->atol_dev) cudaFree(coop->atol_dev);
            coop->atol_dev = nullptr; coop->atol_n = 0;
            e = cudaMalloc((void**)&coop->atol_dev, (size_t)N * sizeof(double));
            if (e != cudaSuccess) return e;
            coop->atol_n = N;
        }
        if (coop->color_bytes < (size_t)N * sizeof(int32_t)) {
            if (coop->color_dev) cudaFree(coop->color_dev);
            coop->color_dev = nullptr; coop->color_bytes = 0;
            e = cudaMalloc(&coop->color_dev, (size_t)N * sizeof(int32_t));
            if (e != cudaSuccess) return e;
            coop->color_bytes = (size_t)N * sizeof(int32_t);
        }
        e = cudaMemcpyAsync(coop->atol_dev, atol_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(coop->color_dev, colmeta.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);              // the host vectors die with this frame
        if (e != cudaSuccess) return e;
        const DsbBandMeta meta{coop->atol_dev, (const int32_t*)coop->color_dev};
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        if constexpr (M::HAS_MASS) {
            // consistent initialisation of the DAE (state.rs:84-162) by the one-lane-per-instance kernel, which hands y, dy,
            // the counters and the status over through the batch-major state arrays
            constexpr int T = DSB_BAND_THREADS_SMALL;
            typedef BandSdirkLayout<M, T> LayS;
            const int64_t want = (pa->nbatch + T - 1) / T;
            const unsigned igrid = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
            const size_t need = (size_t)LayS::WORDS * igrid * T * sizeof(double);
            if (coop->ws_bytes < need) {
                if (coop->ws_mem) cudaFree(coop->ws_mem);
                coop->ws_mem = nullptr; coop->ws_bytes = 0;
                e = cudaMalloc(&coop->ws_mem, need);
                if (e != cudaSuccess) return e;
                coop->ws_bytes = need;
            }
            dsb_band_init_kernel<M, T><<<igrid, T, 0, stream>>>(*pa, *bb, meta, (double*)coop->ws_mem);
            *launches += 1;
        }
        const void* kernel = (const void*)dsb_wband_bdf_solve_dense_kernel<M>;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Lay::THREADS, Lay::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        const int64_t want = (pa->nbatch + Lay::WARPS - 1) / Lay::WARPS;
        const int64_t resident = (int64_t)sms * per_sm;
        const unsigned grid = (unsigned)(want < resident ? want : resident);
        const size_t need = (size_t)Lay::G_WORDS * grid * Lay::WARPS * sizeof(double) + 256;
        if (coop->wb_bytes < need) {
            if (coop->wb_mem) cudaFree(coop->wb_mem);
            coop->wb_mem = nullptr; coop->wb_bytes = 0;
            e = cudaMalloc(&coop->wb_mem, need);
            if (e != cudaSuccess) return e;
            coop->wb_bytes = need;
        }
        // the instance-major result block: the caller's, or our own
        const size_t ys_bytes = (size_t)pa->nt * dsb_model_nout<M>::value * (size_t)pa->nbatch * sizeof(double);
        double* ys_im = coop->ys_im;
        if (!ys_im) {
            if (coop->ys_im_own_bytes < ys_bytes) {
                if (coop->ys_im_own) cudaFree(coop->ys_im_own);
                coop->ys_im_own = nullptr; coop->ys_im_own_bytes = 0;
                e = cudaMalloc((void**)&coop->ys_im_own, ys_bytes);
                if (e != cudaSuccess) return e;
                coop->ys_im_own_bytes = ys_bytes;
            }
            ys_im = coop->ys_im_own;
        }
        e = cudaMemsetAsync(ys_im, 0xFF, ys_bytes, stream);      // outputs never reached stay NaN
        if (e != cudaSuccess) return e;
        if (mid) cudaEventRecord(mid, stream);
        dsb_wband_bdf_solve_dense_kernel<M><<<grid, Lay::THREADS, Lay::SMEM_BYTES, stream>>>(*pa, *bb, meta, (double*)coop->wb_mem, ys_im,
                                                                                              work_counter);
        *launches += 1;
        coop->ys_im_used = ys_im;
        return cudaGetLastError();
    }
};

// lane kernels (one thread per instance): only instantiated for n <= 16
template <class M, bool LANE> struct LaneLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, int, cudaStream_t, cudaEvent_t, unsigned long long*, int*) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct LaneLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, int* launches) {
    const int init_threads = DSB_LANE_THREADS;
    const unsigned init_blocks = (unsigned)((pa->nbatch + init_threads - 1) / init_threads);
    if (method == DSB_METHOD_BDF) {
        const int threads = BdfLayout<M>::THREADS;
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const size_t smem = (size_t)BdfLayout<M>::WORDS * threads * sizeof(double);
        // persistent grid: as many blocks as fit on THIS device at once.  Function attributes and occupancy are per
        // device / context, and a process may drive several GPUs, so both are set and queried on every launch.
        int resident_blocks = 0;
        {
            cudaError_t e = cudaFuncSetAttribute(dsb_bdf_solve_dense_kernel<M>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dsb_bdf_solve_dense_kernel<M>, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            resident_blocks = sms * per_sm;
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        dsb_init_kernel<M><<<init_blocks, init_threads, 0, stream>>>(*pa, *bb, 1);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned grid = blocks < (unsigned)resident_blocks ? blocks : (unsigned)resident_blocks;
        dsb_bdf_solve_dense_kernel<M><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
    } else {
        // (E)SDIRK: the tableau travels in pa->rk; RkState::new_and_consistent uses the tableau order
        const int threads = SdirkLayout<M>::THREADS;
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const size_t smem = (size_t)SdirkLayout<M>::WORDS * threads * sizeof(double);
        int resident_blocks = 0;            // per launch, as above
        {
            cudaError_t e = cudaFuncSetAttribute(dsb_sdirk_solve_dense_kernel<M>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dsb_sdirk_solve_dense_kernel<M>, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            resident_blocks = sms * per_sm;
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        dsb_init_kernel<M><<<init_blocks, init_threads, 0, stream>>>(*pa, *bb, pa->rk.order);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned grid = blocks < (unsigned)resident_blocks ? blocks : (unsigned)resident_blocks;
        dsb_sdirk_solve_dense_kernel<M><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
    }
    return cudaGetLastError();
    }
};

// forward sensitivities (problem.bdf_sens() / tr_bdf2_sens() / esdirk34_sens(), solve_dense_sensitivities): the on-chip lane kernels instantiated for
// DsbWithSens<M> -- ODEs and DAEs of n <= 16 that provide sens_mul / init_sens and have no root / output / reset functions
template <class M, bool LANE> struct SensCapable : std::false_type {};
template <class M> struct SensCapable<M, true>
    : std::bool_constant<dsb_model_has_sens<M>::value && dsb_model_nroots<M>::value == 0 &&
                         !dsb_model_nout<M>::has_out && !dsb_model_has_reset<M>::value> {};
constexpr bool kSensCapable = SensCapable<InstModel, kLaneCapable>::value;
template <class M, bool OK> struct SensLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, int, cudaStream_t, cudaEvent_t, unsigned long long*, DsbCoopState*, int*) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct SensLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, DsbCoopState* coop, int* launches) {
        typedef DsbWithSens<M> MS;
        const bool bdf = method == DSB_METHOD_BDF;
        const int threads = bdf ? BdfLayout<MS>::THREADS : SdirkLayout<MS>::THREADS;
        const size_t smem = (size_t)(bdf ? BdfLayout<MS>::WORDS : SdirkLayout<MS>::WORDS) * threads * sizeof(double);
        const void* kernel = bdf ? (const void*)dsb_bdf_solve_dense_kernel<MS> : (const void*)dsb_sdirk_solve_dense_kernel<MS>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        const unsigned init_blocks = (unsigned)((pa->nbatch + DSB_LANE_THREADS - 1) / DSB_LANE_THREADS);
        dsb_init_kernel<M><<<init_blocks, DSB_LANE_THREADS, 0, stream>>>(*pa, *bb, bdf ? 1 : pa->rk.order);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const unsigned grid = blocks < (unsigned)(sms * per_sm) ? blocks : (unsigned)(sms * per_sm);
        DsbBatchBuffers bbs = *bb;
        bbs.sens_ws = nullptr;
        if (bdf && BdfLayout<MS>::SDIFF_GLOBAL) {        // the sensitivities' difference arrays: one column per resident lane
            const size_t need = (size_t)BdfLayout<MS>::SDIFF_WORDS * grid * threads * sizeof(double);
            if (coop->ws_bytes < need) {
                if (coop->ws_mem) cudaFree(coop->ws_mem);
                coop->ws_mem = nullptr; coop->ws_bytes = 0;
                e = cudaMalloc(&coop->ws_mem, need);
                if (e != cudaSuccess) return e;
                coop->ws_bytes = need;
            }
            bbs.sens_ws = (double*)coop->ws_mem;
        }
        if (bdf) dsb_bdf_solve_dense_kernel<MS><<<grid, threads, smem, stream>>>(*pa, bbs, work_counter);
        else dsb_sdirk_solve_dense_kernel<MS><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
        return cudaGetLastError();
    }
};

// solve(final_time) form (OdeSolverMethod::solve): the on-chip lane kernels instantiated for DsbRagged<M>, equations
// without a reset function
template <class M, bool LANE> struct RaggedCapable : std::false_type {};
template <class M> struct RaggedCapable<M, true> : std::bool_constant<!dsb_model_has_reset<M>::value> {};
constexpr bool kRaggedCapable = RaggedCapable<InstModel, kLaneCapable>::value;
template <class M, bool OK> struct RaggedLauncher {
    static cudaError_t run(const DsbProblemArgs*, const DsbBatchBuffers*, int, cudaStream_t, cudaEvent_t, unsigned long long*, int*) {
        return cudaErrorNotSupported;
    }
};
template <class M> struct RaggedLauncher<M, true> {
    static cudaError_t run(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method, cudaStream_t stream, cudaEvent_t mid,
                           unsigned long long* work_counter, int* launches) {
        typedef DsbRagged<M> MR;
        const bool bdf = method == DSB_METHOD_BDF;
        const int threads = bdf ? BdfLayout<MR>::THREADS : SdirkLayout<MR>::THREADS;
        const size_t smem = (size_t)(bdf ? BdfLayout<MR>::WORDS : SdirkLayout<MR>::WORDS) * threads * sizeof(double);
        const void* kernel = bdf ? (const void*)dsb_bdf_solve_dense_kernel<MR> : (const void*)dsb_sdirk_solve_dense_kernel<MR>;
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        e = cudaMemsetAsync(work_counter, 0, 32 * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        const unsigned init_blocks = (unsigned)((pa->nbatch + DSB_LANE_THREADS - 1) / DSB_LANE_THREADS);
        dsb_init_kernel<M><<<init_blocks, DSB_LANE_THREADS, 0, stream>>>(*pa, *bb, bdf ? 1 : pa->rk.order);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const unsigned grid = blocks < (unsigned)(sms * per_sm) ? blocks : (unsigned)(sms * per_sm);
        if (bdf) dsb_bdf_solve_dense_kernel<MR><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        else dsb_sdirk_solve_dense_kernel<MR><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
        return cudaGetLastError();
    }
};

#if defined(DSB_USER_MODEL_SOURCE)
extern "C"
#endif
cudaError_t DSB_LAUNCH_SYMBOL(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method,
                                                 cudaStream_t stream, cudaEvent_t mid, unsigned long long* work_counter,
                                                 DsbCoopState* coop, const double* atol_host, int* launches) {
    coop->ys_im_used = nullptr;
    // a reset on a DAE (state.apply_reset_with_mass, state.rs:279-306) is built into the on-chip BDF lane kernel only
    if (InstModel::HAS_MASS && dsb_model_has_reset<InstModel>::value &&
        (method != DSB_METHOD_BDF || !kLaneCapable || coop->exec_mode > 1 || pa->ragged || pa->sens))
        return cudaErrorNotSupported;
    if (pa->ragged) {
        if (pa->sens) return cudaErrorNotSupported;
        return RaggedLauncher<InstModel, kRaggedCapable>::run(pa, bb, method, stream, mid, work_counter, launches);
    }
    if (pa->sens) {
        if (!bb->ss) return cudaErrorNotSupported;
        return SensLauncher<InstModel, kSensCapable>::run(pa, bb, method, stream, mid, work_counter, coop, launches);
    }
    // exec_mode 4 / automatic: the warp-per-instance banded kernel (BDF, no reset function)
    if (kWBandCapable && method == DSB_METHOD_BDF && (coop->exec_mode == 4 || (coop->exec_mode == 0 && kWBandPreferred))) {
        const cudaError_t e = WBandLauncher<InstModel, kWBandCapable>::run(pa, bb, stream, mid, work_counter, coop, atol_host, launches);
        if (e != cudaErrorNotSupported || coop->exec_mode == 4) return e;
    } else if (coop->exec_mode == 4) {
        return cudaErrorNotSupported;
    }
    // exec_mode 3 / automatic: the banded lane kernels where the model qualifies (BDF and (E)SDIRK)
    if (kBandCapable && (coop->exec_mode == 3 || coop->exec_mode == 0)) {
        const cudaError_t e = BandLauncher<InstModel, kBandCapable>::run(pa, bb, method, stream, mid, work_counter, coop, atol_host, launches);
        if (e != cudaErrorNotSupported || coop->exec_mode == 3) return e;
    } else if (coop->exec_mode == 3) {
        return cudaErrorNotSupported;
    }
    // root functions (events) and output functions (OdeEquations::out) are built into every kernel family
    // reset functions (re-initialisation after an event) are built into the on-chip lane kernels (BDF and SDIRK) and the
    // block-per-instance kernel, not into the banded lane kernels
    const bool use_coop = coop->exec_mode == 2 || (coop->exec_mode == 0 && !kLaneCapable);
    if (use_coop) {
        // block-per-instance kernel: Bdf, or Sdirk (TR-BDF2 / ESDIRK34)
        return CoopLauncher<InstModel, kCoopBuilt>::run(pa, bb, stream, mid, work_counter, coop, atol_host, launches, method != DSB_METHOD_BDF);
    }
    return LaneLauncher<InstModel, kLaneCapable>::run(pa, bb, method, stream, mid, work_counter, launches);
}

#if defined(DSB_USER_MODEL_SOURCE)
// ---- what a model plugin exports besides its launcher (dsb_capi.cu: dsb_model_library_load) ----------------------------------
extern "C" {
int dsb_plugin_abi(void) { return 2; }
void dsb_plugin_dims(int* n, int* np, int* has_mass, int* nout) {
    *n = InstModel::N; *np = InstModel::NP; *has_mass = InstModel::HAS_MASS ? 1 : 0; *nout = dsb_model_nout<InstModel>::value;
}
// sparsity pattern by NaN probe + greedy colouring with the model's own functor on the host (dsb_host_setup.h: ColoringOf);
// color_full [n], nz_full [n * n] (column-major pattern bytes)
int dsb_plugin_coloring(double t0, DsbProblemArgs* pa, int* probes, int32_t* color_full, uint8_t* nz_full) {
    dsb_problem pr;
    pr.t0 = t0;
    std::vector<int32_t> cf; std::vector<uint8_t> nz;
    dsb_host::ColoringOf f{&pr, pa, probes, &cf, &nz};
    f.template operator()<InstModel>();
    for (size_t k = 0; k < cf.size(); ++k) color_full[k] = cf[k];
    for (size_t k = 0; k < nz.size(); ++k) nz_full[k] = nz[k];
    return (int)cf.size();
}
}
#endif
