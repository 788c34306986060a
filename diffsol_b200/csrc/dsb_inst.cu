// dsb_inst.cu -- instantiates the lane kernels for ONE equation set: compile with -DDSB_INST=<model id>.
#include "dsb_bdf_kernel.cuh"
#include "dsb_init_kernel.cuh"
#include "dsb_launch.h"
#include "dsb_models.h"

#ifndef DSB_INST
#error "compile with -DDSB_INST=<model id>"
#endif
#define DSB_CAT_(a, b) a##b
#define DSB_CAT(a, b) DSB_CAT_(a, b)

typedef dsb_model_by_id<DSB_INST>::type InstModel;

cudaError_t DSB_CAT(dsb_launch_model_, DSB_INST)(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method,
                                                 cudaStream_t stream, cudaEvent_t mid, int* launches) {
    const int threads = DSB_LANE_THREADS;
    const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
    if (method == DSB_METHOD_BDF) {
        const size_t smem = (size_t)BdfLane<InstModel>::SM_WORDS * threads * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(dsb_bdf_solve_dense_kernel<InstModel>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dsb_init_kernel<InstModel><<<blocks, threads, 0, stream>>>(*pa, *bb, 1);
        if (mid) cudaEventRecord(mid, stream);
        dsb_bdf_solve_dense_kernel<InstModel><<<blocks, threads, smem, stream>>>(*pa, *bb);
        *launches += 2;
    } else {
        return cudaErrorNotSupported;
    }
    return cudaGetLastError();
}
