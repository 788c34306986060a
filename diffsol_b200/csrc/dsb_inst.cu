// dsb_inst.cu -- instantiates the lane kernels for ONE equation set: compile with -DDSB_INST=<model id>.
#include "dsb_bdf_kernel.cuh"
#include "dsb_init_kernel.cuh"
#include "dsb_launch.h"
#include "dsb_models.h"
#include "dsb_sdirk_kernel.cuh"

#ifndef DSB_INST
#error "compile with -DDSB_INST=<model id>"
#endif
#define DSB_CAT_(a, b) a##b
#define DSB_CAT(a, b) DSB_CAT_(a, b)

typedef dsb_model_by_id<DSB_INST>::type InstModel;

cudaError_t DSB_CAT(dsb_launch_model_, DSB_INST)(const DsbProblemArgs* pa, const DsbBatchBuffers* bb, int method,
                                                 cudaStream_t stream, cudaEvent_t mid, unsigned long long* work_counter,
                                                 int* launches) {
    const int init_threads = DSB_LANE_THREADS;
    const unsigned init_blocks = (unsigned)((pa->nbatch + init_threads - 1) / init_threads);
    if (method == DSB_METHOD_BDF) {
        const int threads = BdfLayout<InstModel>::THREADS;
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const size_t smem = (size_t)BdfLayout<InstModel>::WORDS * threads * sizeof(double);
        static int resident_blocks = 0;      // persistent grid: as many blocks as fit on the device at once
        if (resident_blocks == 0) {
            cudaError_t e = cudaFuncSetAttribute(dsb_bdf_solve_dense_kernel<InstModel>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dsb_bdf_solve_dense_kernel<InstModel>, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            resident_blocks = sms * per_sm;
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        dsb_init_kernel<InstModel><<<init_blocks, init_threads, 0, stream>>>(*pa, *bb, 1);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned grid = blocks < (unsigned)resident_blocks ? blocks : (unsigned)resident_blocks;
        dsb_bdf_solve_dense_kernel<InstModel><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
    } else {
        // (E)SDIRK: the tableau travels in pa->rk; RkState::new_and_consistent uses the tableau order
        const int threads = SdirkLayout<InstModel>::THREADS;
        const unsigned blocks = (unsigned)((pa->nbatch + threads - 1) / threads);
        const size_t smem = (size_t)SdirkLayout<InstModel>::WORDS * threads * sizeof(double);
        static int resident_blocks = 0;
        if (resident_blocks == 0) {
            cudaError_t e = cudaFuncSetAttribute(dsb_sdirk_solve_dense_kernel<InstModel>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dsb_sdirk_solve_dense_kernel<InstModel>, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            resident_blocks = sms * per_sm;
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        dsb_init_kernel<InstModel><<<init_blocks, init_threads, 0, stream>>>(*pa, *bb, pa->rk.order);
        if (mid) cudaEventRecord(mid, stream);
        const unsigned grid = blocks < (unsigned)resident_blocks ? blocks : (unsigned)resident_blocks;
        dsb_sdirk_solve_dense_kernel<InstModel><<<grid, threads, smem, stream>>>(*pa, *bb, work_counter);
        *launches += 2;
    }
    return cudaGetLastError();
}
