// fp64_peak.cu -- FP64 micro-benchmarks on the box the bench runs on (VERDICT r1 item 2c / SURVEY section 8d: an FP64
// denominator has to be MEASURED before any flop statement).  Prints one JSON object:
//   dfma_tflops            unfused-equivalent peak: 2 flop per DFMA, every SM saturated with independent chains
//   dmma_m8n8k4_tflops     the FP64 tensor-core rate (mma.sync m8n8k4 f64), for the question whether DMMA trailing updates would pay
//   dmul_dadd_tflops       the same with the unfused pair (DMUL then DADD) the bit-exact kernels have to issue (--fmad=false)
//   lat_*_cycles           dependent-issue latency of DFMA / DADD / DMUL, of an IEEE division `a / b`, and of the
//                          reciprocal-reuse division (dsb_math.h: dsb_div_rcp, 5 dependent operations)
//   lds_chain_cycles       dependent shared-memory load (pointer chase)
//   rcp_div_mismatches     dsb_div_rcp(a, b, 1/b) != a / b over `rcp_div_trials` random operand pairs (must be 0)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/_bin/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../diffsol_b200/csrc/dsb_math.h"

template <int MODE>
__global__ void __launch_bounds__(256) throughput_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * 1e-9 + k;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = __fma_rn(a[k], m, c);
            else a[k] = a[k] * m + c;          // --fmad=false: DMUL + DADD
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;
}

// FP64 tensor-core path: mma.sync m8n8k4 f64 (DMMA), 8 independent accumulator pairs per thread.  512 flop per warp
// instruction; every multiply-add of it is FUSED (one rounding), which is why the bit-exact LU cannot use it.
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double c0[8], c1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c0[k] = seed + k; c1[k] = seed - k; }
    const double a = 1.0000001 + threadIdx.x * 1e-12, b = 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c0[k] + c1[k];
    if (s == 12345.678) out[0] = s;
}

// MODE 0 DFMA, 1 DADD, 2 DMUL, 3 IEEE division, 4 reciprocal-reuse division, 5 LDS pointer chase
template <int MODE>
__global__ void latency_kernel(double* out, long long* cycles, int iters, double x0, double d) {
    __shared__ int chase[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) chase[i] = (i + 17) & 255;
    __syncthreads();
    double x = x0;
    const double r = 1.0 / d;
    int p = threadIdx.x & 255;
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) x = __fma_rn(x, d, 1e-9);
        else if (MODE == 1) x = x + d;
        else if (MODE == 2) x = x * d;
        else if (MODE == 3) x = x / d;
        else if (MODE == 4) x = dsb_div_rcp(x, d, r);
        else p = chase[p];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; out[0] = x + p; }
}

__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// random operand pairs, exponents spread over +-300 with a share of adversarial mantissas (all ones, 1 + ulp, equal)
__global__ void rcp_div_check_kernel(unsigned long long* mismatches, int per_thread, uint64_t seed0) {
    uint64_t s = seed0 + 0x1234567ULL * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    unsigned long long bad = 0;
    for (int it = 0; it < per_thread; ++it) {
        uint64_t ma = splitmix(s) & 0x000fffffffffffffULL, mb = splitmix(s) & 0x000fffffffffffffULL;
        const uint64_t e = splitmix(s);
        const int kind = (int)(e >> 60);
        if (kind == 0) mb = 0x000fffffffffffffULL;
        else if (kind == 1) mb = 1;
        else if (kind == 2) ma = 0x000fffffffffffffULL;
        else if (kind == 3) ma = mb;
        else if (kind == 4) mb = 0x000ffffffffffffeULL;
        else if (kind == 5) { ma &= ~0xffffffffULL; mb &= ~0xffffffULL; }
        const uint64_t ea = 1023 - 300 + (e % 601), eb = 1023 - 300 + ((e >> 16) % 601);
        const uint64_t sa = (e >> 40) & 1, sb = (e >> 41) & 1;
        const double a = __longlong_as_double((long long)((sa << 63) | (ea << 52) | ma));
        const double b = __longlong_as_double((long long)((sb << 63) | (eb << 52) | mb));
        const double r = 1.0 / b;
        const double q = dsb_div_rcp(a, b, r), qq = a / b;
        if (__double_as_longlong(q) != __double_as_longlong(qq)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    int dev = 0, sms = 0, clk = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    double* out; long long* cyc; unsigned long long* mism;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&mism, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = sms * 8;
    double tf[2];
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) throughput_kernel<0><<<blocks, 256>>>(out, iters, 1.0);
            else throughput_kernel<1><<<blocks, 256>>>(out, iters, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            const float ms = time_ms(e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        tf[mode] = 2.0 * 8.0 * iters * 256.0 * blocks / (best * 1e-3) / 1e12;
    }
    double dmma_tf = 0.0;
    {
        float best = 1e30f;
        const int diters = 1 << 13;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<<<blocks, 256>>>(out, diters, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            const float ms = time_ms(e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        dmma_tf = 512.0 * 8.0 * diters * (256.0 / 32.0) * blocks / (best * 1e-3) / 1e12;
    }
    double lat[6];
    const int liters = 1 << 14;
    for (int mode = 0; mode < 6; ++mode) {
        long long best = 1ll << 60;
        for (int rep = 0; rep < 4; ++rep) {
            switch (mode) {
                case 0: latency_kernel<0><<<1, 32>>>(out, cyc, liters, 1.0, 0.999999); break;
                case 1: latency_kernel<1><<<1, 32>>>(out, cyc, liters, 1.0, 1e-9); break;
                case 2: latency_kernel<2><<<1, 32>>>(out, cyc, liters, 1.0, 0.9999999); break;
                case 3: latency_kernel<3><<<1, 32>>>(out, cyc, liters, 1.0, 1.0000001); break;
                case 4: latency_kernel<4><<<1, 32>>>(out, cyc, liters, 1.0, 1.0000001); break;
                default: latency_kernel<5><<<1, 32>>>(out, cyc, liters, 1.0, 1.0); break;
            }
            long long c = 0;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            if (c < best) best = c;
        }
        lat[mode] = (double)best / liters;
    }
    cudaMemset(mism, 0, 8);
    const int per_thread = 1 << 12;
    rcp_div_check_kernel<<<sms * 16, 256>>>(mism, per_thread, 0xD1FF501ULL);
    unsigned long long bad = 0;
    cudaMemcpy(&bad, mism, 8, cudaMemcpyDeviceToHost);
    const cudaError_t err = cudaGetLastError();
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"dfma_tflops\": %.3f, \"dmul_dadd_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
           "\"lat_dfma_cycles\": %.2f, \"lat_dadd_cycles\": %.2f, \"lat_dmul_cycles\": %.2f, \"lat_ieee_div_cycles\": %.2f, "
           "\"lat_rcp_div_cycles\": %.2f, \"lds_chain_cycles\": %.2f, \"rcp_div_trials\": %llu, \"rcp_div_mismatches\": %llu, "
           "\"cuda_error\": \"%s\"}\n",
           prop.name, sms, clk, tf[0], tf[1], dmma_tf, lat[0], lat[1], lat[2], lat[3], lat[4], lat[5],
           (unsigned long long)sms * 16ull * 256ull * per_thread, bad, cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
