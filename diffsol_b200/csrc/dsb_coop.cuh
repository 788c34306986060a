// dsb_coop.cuh -- block-cooperative (one thread block per instance) building blocks for systems too large
// for the one-thread-per-instance kernels (n > 16): dense LU with partial pivoting and the two triangular
// solves, with the matrix in global memory (column-major, leading dimension n: 512 KB at n = 256 does not
// fit in shared memory) and panels staged through shared memory.
//
// Arithmetic is the nalgebra 0.35 LU the reference calls (diffsol-la/src/linear_solver/nalgebra/lu.rs:31-51):
// first maximum as pivot, reciprocal-pivot scaling of the sub-column, rank-1 updates `a = (-u) * l + a`
// applied to every element in ascending pivot order, column-axpy substitutions.  The factorisation is
// BLOCKED (panel width 32: the trailing matrix is streamed through the SM once per panel instead of once
// per column) but every element still receives exactly the same sequence of unfused multiply / add
// operations as in the unblocked right-looking algorithm, so factors, pivots and solutions are bit-identical
// to the CPU path.  (Deferring a panel's row swaps for the columns outside the panel moves elements before
// they are updated instead of after; the multiplier and the pivot-row entry an element meets at step i are
// the same either way.)  FP64 tensor-core MMA is deliberately NOT used for the trailing update: DMMA fuses
// the multiply-add and would change the rounding of every element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsb_math.h"

#define DSB_COOP_NB 32            // panel width
#define DSB_COOP_MAX_N 512

// Shared-memory scratch of one block: a panel of up to n x NB doubles + reduction scratch.
struct CoopScratch {
    double* panel;     // [NB][m] column-major, m = rows below and including the panel's first row
    double* redv;      // [32] per-warp partial maxima
    int* redi;         // [32]
    int* bcast;        // [4] broadcast words
};

__device__ __forceinline__ size_t coop_lu_smem_bytes(int n) {
    return (size_t)n * DSB_COOP_NB * sizeof(double) + 32 * sizeof(double) + 32 * sizeof(int) + 4 * sizeof(int);
}
inline size_t coop_lu_smem_bytes_host(int n) {
    return (size_t)n * DSB_COOP_NB * sizeof(double) + 32 * sizeof(double) + 32 * sizeof(int) + 4 * sizeof(int);
}

__device__ __forceinline__ CoopScratch coop_carve(void* smem, int n) {
    CoopScratch s;
    s.panel = (double*)smem;
    s.redv = s.panel + (size_t)n * DSB_COOP_NB;
    s.redi = (int*)(s.redv + 32);
    s.bcast = s.redi + 32;
    return s;
}

// In-place LU of the n x n column-major matrix A (global memory, leading dimension n) by the whole block.
// piv[i] = row swapped with row i (== i: no swap).  Returns (to every thread) 0, or k+1 for the first
// zero pivot column k (nalgebra leaves that column untouched and continues).
static __device__ __noinline__ int coop_lu_factor(double* __restrict__ A, int n, int* __restrict__ piv, const CoopScratch& sc) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = (T + 31) >> 5;
    int first_bad = 0;
    for (int k0 = 0; k0 < n; k0 += DSB_COOP_NB) {
        const int kb = (n - k0 < DSB_COOP_NB) ? (n - k0) : DSB_COOP_NB;
        const int m = n - k0;                        // panel rows
        double* P = sc.panel;                        // P[c * m + lr]
        // ---- 1. stage the panel ----
        for (int c = 0; c < kb; ++c)
            for (int lr = tid; lr < m; lr += T) P[(size_t)c * m + lr] = A[(size_t)(k0 + c) * n + k0 + lr];
        __syncthreads();
        // ---- 2. factor the panel (unblocked, in shared memory) ----
        for (int i = 0; i < kb; ++i) {
            // pivot = first maximum of |P[i][lr]|, lr >= i (NaNs below the diagonal never win; a NaN on the
            // diagonal keeps the diagonal, as `val > the_max` is false for every candidate)
            double bv = -1.0; int bi = 0x7fffffff;
            for (int lr = i + tid; lr < m; lr += T) {
                const double v = dsb_abs(P[(size_t)i * m + lr]);
                if (v > bv) { bv = v; bi = lr; }        // ascending lr per thread: keeps the first maximum
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { sc.redv[wid] = bv; sc.redi[wid] = bi; }
            __syncthreads();
            if (wid == 0) {
                bv = (lane < nwarps) ? sc.redv[lane] : -1.0;
                bi = (lane < nwarps) ? sc.redi[lane] : 0x7fffffff;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) {
                    const double dii = P[(size_t)i * m + i];
                    int p = bi;
                    if (dii != dii) p = i;                             // NaN diagonal: the_max = NaN, never replaced
                    const double diag = P[(size_t)i * m + p];
                    const int ok = (diag == 0.0) ? 0 : 1;
                    if (!ok) p = i;
                    sc.bcast[0] = p; sc.bcast[1] = ok;
                    piv[k0 + i] = k0 + p;
                }
            }
            __syncthreads();
            const int p = sc.bcast[0];
            const int ok = sc.bcast[1];
            if (!ok) { if (first_bad == 0) first_bad = k0 + i + 1; __syncthreads(); continue; }
            if (p != i) {
                for (int c = tid; c < kb; c += T) {
                    const double tmp = P[(size_t)c * m + i]; P[(size_t)c * m + i] = P[(size_t)c * m + p]; P[(size_t)c * m + p] = tmp;
                }
            }
            __syncthreads();
            const double inv_diag = 1.0 / P[(size_t)i * m + i];
            for (int lr = i + 1 + tid; lr < m; lr += T) P[(size_t)i * m + lr] *= inv_diag;
            __syncthreads();
            // rank-1 update of the remaining panel columns: a thread owns rows, walks the columns
            for (int lr = i + 1 + tid; lr < m; lr += T) {
                const double l = P[(size_t)i * m + lr];
                for (int c = i + 1; c < kb; ++c) {
                    const double mpk = -P[(size_t)c * m + i];
                    P[(size_t)c * m + lr] = mpk * l + P[(size_t)c * m + lr];
                }
            }
            __syncthreads();
        }
        // ---- 3. write the panel back ----
        for (int c = 0; c < kb; ++c)
            for (int lr = tid; lr < m; lr += T) A[(size_t)(k0 + c) * n + k0 + lr] = P[(size_t)c * m + lr];
        // ---- 4. row swaps for the columns outside the panel; U12 = L11^-1 A12 (one thread per column, the
        //         kb-long column segment held in registers) ----
        for (int c = tid; c < n; c += T) {
            if (c >= k0 && c < k0 + kb) continue;
            double* col = A + (size_t)c * n;
            for (int i = 0; i < kb; ++i) {
                const int p = piv[k0 + i];
                if (p != k0 + i) { const double tmp = col[k0 + i]; col[k0 + i] = col[p]; col[p] = tmp; }
            }
            if (c >= k0 + kb) {
                if (kb == DSB_COOP_NB) {
                    double seg[DSB_COOP_NB];
#pragma unroll
                    for (int r = 0; r < DSB_COOP_NB; ++r) seg[r] = col[k0 + r];
#pragma unroll
                    for (int i = 0; i < DSB_COOP_NB; ++i) {
                        const double mpk = -seg[i];
#pragma unroll
                        for (int r = i + 1; r < DSB_COOP_NB; ++r) seg[r] = mpk * P[(size_t)i * m + r] + seg[r];
                    }
#pragma unroll
                    for (int r = 0; r < DSB_COOP_NB; ++r) col[k0 + r] = seg[r];
                } else {
                    for (int i = 0; i < kb; ++i) {
                        const double mpk = -col[k0 + i];
                        for (int r = i + 1; r < kb; ++r) col[k0 + r] = mpk * P[(size_t)i * m + r] + col[k0 + r];
                    }
                }
            }
        }
        __syncthreads();
        // ---- 5. trailing update A22[r][c] = sum_i (-U[i][c]) * L[r][i] + A22[r][c], i ascending ----
        // warps own groups of 4 columns (4 independent accumulation chains per lane); lanes own rows in chunks
        // of 32; L21 rows come from the staged panel, U12 entries are warp-uniform (broadcast) loads
        const int r0 = k0 + kb;
        const int m2 = n - r0;
        if (m2 > 0) {
            for (int rc = 0; rc < m2; rc += 32) {
                const int r = r0 + rc + lane;
                const bool live = r < n;
                double l[DSB_COOP_NB];
#pragma unroll
                for (int i = 0; i < DSB_COOP_NB; ++i) l[i] = (live && i < kb) ? P[(size_t)i * m + (r - k0)] : 0.0;
                for (int c = r0 + 4 * wid; c < n; c += 4 * nwarps) {
                    const int nc = (n - c < 4) ? (n - c) : 4;
                    double* col0 = A + (size_t)c * n;
                    double* col1 = A + (size_t)(c + (nc > 1 ? 1 : 0)) * n;
                    double* col2 = A + (size_t)(c + (nc > 2 ? 2 : 0)) * n;
                    double* col3 = A + (size_t)(c + (nc > 3 ? 3 : 0)) * n;
                    double a0 = live ? col0[r] : 0.0, a1 = live ? col1[r] : 0.0, a2 = live ? col2[r] : 0.0, a3 = live ? col3[r] : 0.0;
                    if (kb == DSB_COOP_NB) {
#pragma unroll
                        for (int i = 0; i < DSB_COOP_NB; ++i) {
                            const double li = l[i];
                            a0 = (-col0[k0 + i]) * li + a0;
                            a1 = (-col1[k0 + i]) * li + a1;
                            a2 = (-col2[k0 + i]) * li + a2;
                            a3 = (-col3[k0 + i]) * li + a3;
                        }
                    } else {
                        for (int i = 0; i < kb; ++i) {
                            const double li = l[i];
                            a0 = (-col0[k0 + i]) * li + a0;
                            a1 = (-col1[k0 + i]) * li + a1;
                            a2 = (-col2[k0 + i]) * li + a2;
                            a3 = (-col3[k0 + i]) * li + a3;
                        }
                    }
                    if (live) {
                        col0[r] = a0;
                        if (nc > 1) col1[r] = a1;
                        if (nc > 2) col2[r] = a2;
                        if (nc > 3) col3[r] = a3;
                    }
                }
            }
        }
        __syncthreads();
    }
    return first_bad;
}

// Solve with the factors of coop_lu_factor; b lives in SHARED memory (n doubles).  Returns false (to every
// thread) when U has a zero on its diagonal (nalgebra's solve_mut returns false; b is then unspecified).
static __device__ __noinline__ bool coop_lu_solve(const double* __restrict__ LU, int n, const int* __restrict__ piv, double* b,
                              const CoopScratch& sc) {
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) {
        for (int i = 0; i < n; ++i) { const int p = piv[i]; if (p != i) { const double tmp = b[i]; b[i] = b[p]; b[p] = tmp; } }
        sc.bcast[2] = 1;
    }
    __syncthreads();
    // forward substitution with the unit lower triangle, column-axpy form, in blocks of 32 columns.  Every
    // thread first loads the (up to) 32 factor entries its row needs (independent loads in flight together),
    // then runs the dependent multiply-add chain out of registers.
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int kb = (n - k0 < 32) ? (n - k0) : 32;
        double lv[32];
        const int rb = k0 + kb + tid;                     // this thread's row below the block (if any)
        if (tid < 32) {                                   // diagonal block: one warp, lane = row within the block
            const int r = k0 + tid;
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb && tid < kb) ? LU[(size_t)(k0 + i) * n + r] : 0.0;
            double br = (tid < kb) ? b[r] : 0.0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const double coeff = __shfl_sync(0xffffffffu, br, i);
                if (i < kb && tid > i && tid < kb) br = (-coeff) * lv[i] + br;
            }
            if (tid < kb) b[r] = br;
        }
        // (rows below are one per thread for n <= 32 + blockDim; the generic loop handles larger n)
        if (rb < n) {
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb) ? LU[(size_t)(k0 + i) * n + rb] : 0.0;
        }
        __syncthreads();
        if (rb < n) {
            double br = b[rb];
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < kb) br = (-b[k0 + i]) * lv[i] + br;
            b[rb] = br;
        }
        for (int r = rb + T; r < n; r += T) {
            double br = b[r];
            for (int i = 0; i < kb; ++i) br = (-b[k0 + i]) * LU[(size_t)(k0 + i) * n + r] + br;
            b[r] = br;
        }
        __syncthreads();
    }
    // back substitution with U, column-axpy form, blocks of 32 columns from the bottom
    const int nblk = (n + 31) / 32;
    for (int kbk = nblk - 1; kbk >= 0; --kbk) {
        const int k0 = kbk * 32;
        const int kb = (n - k0 < 32) ? (n - k0) : 32;
        double lv[32];
        if (tid < 32) {
            const int r = k0 + tid;
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb && tid < kb) ? LU[(size_t)(k0 + i) * n + r] : 0.0;
            double br = (tid < kb) ? b[r] : 0.0;
            bool ok = true;
#pragma unroll
            for (int i = 31; i >= 0; --i) {
                if (i < kb && ok) {
                    const double diag = __shfl_sync(0xffffffffu, lv[i], i);       // U[i][i] lives in lane i
                    if (diag == 0.0) ok = false;                                    // uniform across the warp
                    else {
                        double coeff = 0.0;
                        if (tid == i) { coeff = br / diag; br = coeff; }
                        coeff = __shfl_sync(0xffffffffu, coeff, i);
                        if (tid < i) br = (-coeff) * lv[i] + br;
                    }
                }
            }
            if (tid < kb) b[r] = br;
            if (!ok && tid == 0) sc.bcast[2] = 0;
        }
        if (tid < k0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb) ? LU[(size_t)(k0 + i) * n + tid] : 0.0;
        }
        __syncthreads();
        if (sc.bcast[2] == 0) return false;
        if (tid < k0) {                                    // rows above: contributions in DESCENDING i
            double br = b[tid];
#pragma unroll
            for (int i = 31; i >= 0; --i) if (i < kb) br = (-b[k0 + i]) * lv[i] + br;
            b[tid] = br;
        }
        for (int r = tid + T; r < k0; r += T) {
            double br = b[r];
            for (int i = kb - 1; i >= 0; --i) br = (-b[k0 + i]) * LU[(size_t)(k0 + i) * n + r] + br;
            b[r] = br;
        }
        __syncthreads();
    }
    return true;
}


// ---- banded variant -------------------------------------------------------------------------------------------
// PDE-type systems (BASELINE configs 4 and 5) have banded Jacobians: almost every multiplier and pivot-row
// entry of the dense elimination is an exact zero, and `a = (-0) * l + a` leaves `a` untouched.  When the
// iteration matrix has lower / upper bandwidths kl / ku with 2 kl + ku + 1 <= 32 it is factored in LAPACK band
// storage held in SHARED memory (ab[j * 32 + kv + r - j], kv = kl + ku: room for the fill-in of partial
// pivoting), by the first warp of the block, at O(n kl (kl + ku)) work, and the substitutions of every Newton
// iteration read the band at shared-memory latency: O(n (2 kl + ku)) instead of O(n^2) bytes from HBM.
//
// Same arithmetic as the dense algorithm: first maximum as pivot, reciprocal-pivot scaling, updates
// `a = (-u) * l + a` in ascending pivot order, column-axpy substitutions; every operation that is skipped has an
// exactly zero multiplier or pivot-row entry.  Unlike nalgebra, the row interchanges are not applied to the
// multipliers of earlier columns (the dgbtf2 convention); the forward substitution therefore interleaves the
// interchanges with the column updates.  Every right-hand-side entry still meets the same multipliers in the
// same order, so the solutions are bit-identical to nalgebra's.

// kl, ku of the union of the non-zero patterns of J and (optionally) M; every thread of the block takes part.
static __device__ __noinline__ void coop_band_scan(const double* __restrict__ J, const double* __restrict__ Mm, int n, int* kl_ku) {
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) { kl_ku[0] = 0; kl_ku[1] = 0; }
    __syncthreads();
    int kl = 0, ku = 0;
    for (int r = tid; r < n; r += T) {                    // a thread owns a row: loads are coalesced across the warp
        for (int c = 0; c < n; ++c) {
            const bool nz = J[(size_t)c * n + r] != 0.0 || (Mm != nullptr && Mm[(size_t)c * n + r] != 0.0);
            if (nz) { if (r - c > kl) kl = r - c; if (c - r > ku) ku = c - r; }
        }
    }
    atomicMax(&kl_ku[0], kl);
    atomicMax(&kl_ku[1], ku);
    __syncthreads();
}

// dgbtf2-style band LU in shared memory; executed by warp 0 only.  piv[j] = row interchanged with row j.
static __device__ __noinline__ int warp_band_factor(double* ab, int n, int kl, int ku, int* __restrict__ piv) {
    const int lane = threadIdx.x & 31;
    const int kv = kl + ku;
    int first_bad = 0;
    int ju = 0;                                            // last column touched by the fill-in so far
    for (int j = 0; j < n; ++j) {
        const int km = (kl < n - 1 - j) ? kl : (n - 1 - j);
        // pivot: first maximum of |A[j + d][j]|, d = 0 .. km
        double bv = -1.0; int bi = 0x7fffffff;
        if (lane <= km) { bv = dsb_abs(ab[j * 32 + kv + lane]); bi = lane; if (!(bv == bv)) { bv = -1.0; bi = 0x7fffffff; } }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        const double dii = ab[j * 32 + kv];
        int jp = (bi == 0x7fffffff) ? 0 : bi;
        if (dii != dii) jp = 0;                            // NaN diagonal keeps the diagonal
        const double diag = ab[j * 32 + kv + jp];
        if (diag == 0.0) { if (lane == 0) piv[j] = j; if (first_bad == 0) first_bad = j + 1; __syncwarp(); continue; }
        if (lane == 0) piv[j] = j + jp;
        { const int cand = (j + ku + jp < n - 1) ? (j + ku + jp) : (n - 1); if (cand > ju) ju = cand; }
        if (jp != 0) {
            const int c = j + lane;                         // columns j .. ju (at most kv + 1 <= 32 of them)
            if (c <= ju) {
                const double a = ab[c * 32 + kv + j - c], b = ab[c * 32 + kv + j + jp - c];
                ab[c * 32 + kv + j - c] = b; ab[c * 32 + kv + j + jp - c] = a;
            }
            __syncwarp();
        }
        if (km > 0) {
            const double inv_diag = 1.0 / ab[j * 32 + kv];
            if (lane >= 1 && lane <= km) ab[j * 32 + kv + lane] *= inv_diag;
            __syncwarp();
            const int ncols = ju - j;                       // columns j+1 .. ju
            for (int e = lane; e < ncols * km; e += 32) {
                const int c = j + 1 + e / km, d = 1 + e % km;       // row j + d
                const double mpk = -ab[c * 32 + kv + j - c];
                ab[c * 32 + kv + j + d - c] = mpk * ab[j * 32 + kv + d] + ab[c * 32 + kv + j + d - c];
            }
            __syncwarp();
        }
    }
    return first_bad;
}

// dgbtrs-style substitutions on the shared-memory band; executed by warp 0 only; b in shared memory.
static __device__ __noinline__ bool warp_band_solve(const double* ab, int n, int kl, int ku, const int* __restrict__ piv, double* b) {
    const int lane = threadIdx.x & 31;
    const int kv = kl + ku;
    if (kl > 0) {
        for (int j = 0; j + 1 < n; ++j) {
            const int lm = (kl < n - 1 - j) ? kl : (n - 1 - j);
            const int l = piv[j];
            if (l != j) { __syncwarp(); if (lane == 0) { const double tmp = b[l]; b[l] = b[j]; b[j] = tmp; } __syncwarp(); }
            const double mc = -b[j];
            if (lane >= 1 && lane <= lm) b[j + lane] = mc * ab[j * 32 + kv + lane] + b[j + lane];
            __syncwarp();
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        const double diag = ab[i * 32 + kv];
        if (diag == 0.0) return false;
        const double coeff = b[i] / diag;
        __syncwarp();
        if (lane == 0) b[i] = coeff;
        const int r = i - 1 - lane;                         // lane e - 1 handles U[i - e][i], e = 1 .. kv
        if (lane < kv && r >= 0) b[r] = (-coeff) * ab[i * 32 + kv - 1 - lane] + b[r];
        __syncwarp();
    }
    return true;
}

// ---- the same band routines for ONE thread, kl / ku known at compile time --------------------------------------------
// With a narrow band (PDE stencils: kl = ku = 1) a warp has nothing to share out: 2 or 3 of 32 lanes work, and every
// column costs two warp synchronisations.  One thread walking the band with its running entries in a REGISTER WINDOW
// (kl + 1 values in the forward pass, kl + ku + 1 in the backward pass) and the pivots in SHARED memory is ~10x
// faster per substitution at n = 256: the loads of band entries and pivots do not depend on the recurrence and run
// ahead of it.  Same element-wise arithmetic and order as warp_band_factor / warp_band_solve above (and hence as
// nalgebra's dense LU).  piv[j] = row interchanged with row j.
template <int KL, int KU>
static __device__ __noinline__ int thread_band_factor(double* ab, int n, int* piv) {
    constexpr int KV = KL + KU;
    int first_bad = 0;
    int ju = 0;
    for (int j = 0; j < n; ++j) {
        const int km = (KL < n - 1 - j) ? KL : (n - 1 - j);
        double colv[KL + 1];
#pragma unroll
        for (int d = 0; d <= KL; ++d) colv[d] = (d <= km) ? ab[j * 32 + KV + d] : 0.0;
        int jp = 0;
        double best = -1.0;
#pragma unroll
        for (int d = 0; d <= KL; ++d) {
            const double av = dsb_abs(colv[d]);
            if (d <= km && av == av && av > best) { best = av; jp = d; }
        }
        if (colv[0] != colv[0]) jp = 0;
        double diag = colv[0];
#pragma unroll
        for (int d = 1; d <= KL; ++d) if (jp == d) diag = colv[d];
        if (diag == 0.0) { piv[j] = j; if (first_bad == 0) first_bad = j + 1; continue; }
        piv[j] = j + jp;
        { const int cand = (j + KU + jp < n - 1) ? (j + KU + jp) : (n - 1); if (cand > ju) ju = cand; }
        if (jp != 0) {
#pragma unroll
            for (int q = 0; q <= KV; ++q) {
                const int c = j + q;
                if (c <= ju) {
                    const double a = ab[c * 32 + KV - q], b = ab[c * 32 + KV - q + jp];
                    ab[c * 32 + KV - q] = b; ab[c * 32 + KV - q + jp] = a;
                }
            }
#pragma unroll
            for (int d = 1; d <= KL; ++d) if (jp == d) { const double a = colv[0]; colv[0] = colv[d]; colv[d] = a; }
        }
        if (km > 0) {
            const double inv_diag = 1.0 / colv[0];
#pragma unroll
            for (int d = 1; d <= KL; ++d) if (d <= km) { colv[d] *= inv_diag; ab[j * 32 + KV + d] = colv[d]; }
#pragma unroll
            for (int q = 1; q <= KV; ++q) {
                const int c = j + q;
                if (c <= ju) {
                    const double mpk = -ab[c * 32 + KV - q];
#pragma unroll
                    for (int d = 1; d <= KL; ++d)
                        if (d <= km) ab[c * 32 + KV - q + d] = mpk * colv[d] + ab[c * 32 + KV - q + d];
                }
            }
        }
    }
    return first_bad;
}

template <int KL, int KU>
static __device__ __noinline__ bool thread_band_solve(const double* ab, int n, const int* piv, double* b) {
    constexpr int KV = KL + KU;
    {
        double w[KL + 1];
#pragma unroll
        for (int d = 0; d <= KL; ++d) w[d] = (d < n) ? b[d] : 0.0;
#pragma unroll 4
        for (int j = 0; j + 1 < n; ++j) {
            const int jp = piv[j] - j;
            if (jp != 0) {
                const double a = w[0];
#pragma unroll
                for (int d = 1; d <= KL; ++d) if (jp == d) { w[0] = w[d]; w[d] = a; }
            }
            const double bj = w[0];
            b[j] = bj;
            const double nbj = -bj;
            const int lm = (KL < n - 1 - j) ? KL : (n - 1 - j);
#pragma unroll
            for (int d = 1; d <= KL; ++d) if (d <= lm) w[d] = nbj * ab[j * 32 + KV + d] + w[d];
#pragma unroll
            for (int d = 0; d < KL; ++d) w[d] = w[d + 1];
            w[KL] = (j + 1 + KL < n) ? b[j + 1 + KL] : 0.0;
        }
        b[n - 1] = w[0];
    }
    double w[KV + 1];
#pragma unroll
    for (int e = 0; e <= KV; ++e) w[e] = (n - 1 - e >= 0) ? b[n - 1 - e] : 0.0;
#pragma unroll 4
    for (int i = n - 1; i >= 0; --i) {
        const double diag = ab[i * 32 + KV];
        if (diag == 0.0) return false;
        const double coeff = w[0] / diag;
        b[i] = coeff;
        const double ncoeff = -coeff;
#pragma unroll
        for (int e = 1; e <= KV; ++e) if (i - e >= 0) w[e] = ncoeff * ab[i * 32 + KV - e] + w[e];
#pragma unroll
        for (int e = 0; e < KV; ++e) w[e] = w[e + 1];
        w[KV] = (i - 1 - KV >= 0) ? b[i - 1 - KV] : 0.0;
    }
    return true;
}

