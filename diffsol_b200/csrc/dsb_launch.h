// dsb_launch.h -- one launcher per built-in equation set; each lives in its own translation unit
// (dsb_inst.cu compiled with -DDSB_INST=<model id>) so the heavily unrolled lane kernels build in parallel.
#pragma once
#include <cuda_runtime.h>

#include "dsb_args.h"

// `mid` (may be NULL) is recorded between the initialisation kernel and the integrator kernel so the
// integrator's own duration can be read back; `work_counter` is the device word the persistent integrator
// kernel draws instance indices from (zeroed by the launcher).
// `coop` describes the block-per-instance path: exec_mode (0 = automatic: lane kernels for n <= 16, banded kernels for
// banded models above that -- warp per instance for BDF, lane per instance otherwise --, cooperative kernel for the
// rest; 1 = lane; 2 = cooperative; 3 = lane per instance, banded, state in global memory; 4 = warp per instance,
// banded, state in shared memory) and its global-memory workspace (grown on demand by the launcher).
struct DsbCoopState {
    int exec_mode;
    void* ws_mem; size_t ws_bytes;
    double* atol_dev; int atol_n;
    const int32_t* color_host; const uint8_t* nz_host;   // colouring of the current solve (host, valid during the launch call)
    void* color_dev; size_t color_bytes;                 // device copy: [n] int32 colours then [n*n] pattern bytes
    // warp-per-instance banded kernel (dsb_wband_bdf_kernel.cuh): per-warp global slots (df/dy and M bands) and the
    // INSTANCE-major result block it writes.  ys_im: [B][nt][nout] device buffer offered by the caller (may be NULL: the
    // launcher then grows ys_im_own); ys_im_used: set by the launcher to the buffer the kernel wrote, NULL when the
    // kernel that ran wrote the batch-major `bb.ys` instead.
    void* wb_mem; size_t wb_bytes;
    double* ys_im; double* ys_im_own; size_t ys_im_own_bytes; double* ys_im_used;
};
typedef cudaError_t (*dsb_launch_fn)(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method,
                                     cudaStream_t stream, cudaEvent_t mid, unsigned long long* work_counter,
                                     DsbCoopState* coop, const double* atol_host, int* launches);

#define DSB_DECLARE_LAUNCH(id) cudaError_t dsb_launch_model_##id(const DsbProblemArgs*, const DsbBatchBuffers*, int, cudaStream_t, cudaEvent_t, unsigned long long*, DsbCoopState*, const double*, int*);
DSB_DECLARE_LAUNCH(0) DSB_DECLARE_LAUNCH(1) DSB_DECLARE_LAUNCH(2) DSB_DECLARE_LAUNCH(3)
DSB_DECLARE_LAUNCH(4) DSB_DECLARE_LAUNCH(5) DSB_DECLARE_LAUNCH(6) DSB_DECLARE_LAUNCH(7)
DSB_DECLARE_LAUNCH(8) DSB_DECLARE_LAUNCH(9) DSB_DECLARE_LAUNCH(10) DSB_DECLARE_LAUNCH(11)
DSB_DECLARE_LAUNCH(12) DSB_DECLARE_LAUNCH(13) DSB_DECLARE_LAUNCH(14) DSB_DECLARE_LAUNCH(15)
DSB_DECLARE_LAUNCH(16) DSB_DECLARE_LAUNCH(17) DSB_DECLARE_LAUNCH(18) DSB_DECLARE_LAUNCH(19) DSB_DECLARE_LAUNCH(20) DSB_DECLARE_LAUNCH(21) DSB_DECLARE_LAUNCH(22)
#undef DSB_DECLARE_LAUNCH
