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
#define DSB_COOP_UT_WORDS (2 * DSB_COOP_NB * DSB_COOP_NB)     // two 32 x 32 tiles of U12 (double-buffered), see the trailing update
struct CoopScratch {
    double* panel;     // [NB][m] column-major, m = rows below and including the panel's first row
    double* utile;     // [2][NB columns][NB rows]: tiles of U12 for the trailing update
    double* redv;      // [32] per-warp partial maxima
    int* redi;         // [32]
    int* bcast;        // [4] broadcast words
};

__device__ __forceinline__ size_t coop_lu_smem_bytes(int n) {
    return ((size_t)n * DSB_COOP_NB + DSB_COOP_UT_WORDS) * sizeof(double) + 32 * sizeof(double) + 32 * sizeof(int) + 4 * sizeof(int);
}
inline size_t coop_lu_smem_bytes_host(int n) {
    return ((size_t)n * DSB_COOP_NB + DSB_COOP_UT_WORDS) * sizeof(double) + 32 * sizeof(double) + 32 * sizeof(int) + 4 * sizeof(int);
}

__device__ __forceinline__ CoopScratch coop_carve(void* smem, int n) {
    CoopScratch s;
    s.panel = (double*)smem;
    s.utile = s.panel + (size_t)n * DSB_COOP_NB;
    s.redv = s.utile + DSB_COOP_UT_WORDS;
    s.redi = (int*)(s.redv + 32);
    s.bcast = s.redi + 32;
    return s;
}

// ---- bulk-async copies (1-D TMA, cp.async.bulk) with mbarrier completion -----------------------------------------------
// The panels of the factorisation and of the triangular solves are runs of contiguous column segments of the
// column-major matrix: one thread issues one bulk copy per column segment, all completing on one mbarrier, and the
// block waits on that barrier -- no thread spends registers or issue slots on the staging, and the copy of the NEXT
// panel runs under the arithmetic of the current one.  Segment addresses and lengths must be multiples of 16 bytes:
// n even and panel origins at multiples of 16 rows, else the callers fall back to plain loads.
// The barriers live in the tail of the reduction scratch (redv[28..31]; the pivot search uses redv[warp], <= 4 warps).
#define DSB_COOP_BAR(sc, k) ((sc).redv + 28 + (k))
__device__ __forceinline__ void coop_mbar_init(double* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void coop_mbar_inval(double* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void coop_mbar_expect(double* bar, unsigned bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async;" ::: "memory");                 // earlier generic-proxy accesses to the buffers
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void coop_bulk_load(double* sdst, const double* gsrc, unsigned bytes, double* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void coop_bulk_store(double* gdst, const double* ssrc, unsigned bytes) {
    const unsigned s_ = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s_), "r"(bytes) : "memory");
}
__device__ __forceinline__ void coop_mbar_wait(double* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}

// In-place LU of the n x n column-major matrix A (global memory, leading dimension n) by the whole block.
// piv[i] = row swapped with row i (== i: no swap).  Returns (to every thread) 0, or k+1 for the first
// zero pivot column k (nalgebra leaves that column untouched and continues).
static __device__ __noinline__ int coop_lu_factor(double* __restrict__ A, int n, int* __restrict__ piv, const CoopScratch& sc) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nwarps = (T + 31) >> 5;
    int first_bad = 0;
#ifndef DSB_COOP_NO_BULK
    const bool bulk = (n & 1) == 0 && ((uintptr_t)A & 15) == 0;   // 16-byte alignment of every column segment
#else
    const bool bulk = false;
#endif
    double* const fbar = DSB_COOP_BAR(sc, 2);
    if (bulk) {
        if (tid == 0) coop_mbar_init(fbar);
        __syncthreads();
    }
    unsigned fparity = 0;
    for (int k0 = 0; k0 < n; k0 += DSB_COOP_NB) {
        const int kb = (n - k0 < DSB_COOP_NB) ? (n - k0) : DSB_COOP_NB;
        const int m = n - k0;                        // panel rows
        double* P = sc.panel;                        // P[c * m + lr]
        // ---- 1. stage the panel: one bulk copy per column segment (plain loads when they cannot be aligned) ----
        if (bulk) {
            if (tid == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous panel's write-back has left P
                coop_mbar_expect(fbar, (unsigned)(kb * m * 8));
                for (int c = 0; c < kb; ++c) coop_bulk_load(P + (size_t)c * m, A + (size_t)(k0 + c) * n + k0, (unsigned)(m * 8), fbar);
            }
            coop_mbar_wait(fbar, fparity);
            fparity ^= 1u;
        } else {
            for (int c = 0; c < kb; ++c)
                for (int lr = tid; lr < m; lr += T) P[(size_t)c * m + lr] = A[(size_t)(k0 + c) * n + k0 + lr];
            __syncthreads();
        }
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
        // ---- 3. write the panel back (bulk: asynchronous; nothing below reads these columns of A, P stays valid) ----
        if (bulk) {
            asm volatile("fence.proxy.async;" ::: "memory");       // every thread's writes to P, before the async proxy reads it
            __syncthreads();
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the block's writes to P (ordered by the barrier above)
                for (int c = 0; c < kb; ++c) coop_bulk_store(A + (size_t)(k0 + c) * n + k0, P + (size_t)c * m, (unsigned)(m * 8));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int c = 0; c < kb; ++c)
                for (int lr = tid; lr < m; lr += T) A[(size_t)(k0 + c) * n + k0 + lr] = P[(size_t)c * m + lr];
        }
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
        // Register-tiled: a lane owns TWO rows (r, r + 32) of FOUR columns -- 8 independent accumulation chains -- and walks
        // i = 0 .. 31 once; L21 comes from the staged panel (conflict-free: consecutive lanes, consecutive rows), U12 from a
        // 32 x 32 tile in shared memory read as warp-wide broadcasts, two i at a time (16-byte loads).  The tiles of U12 --
        // 32 column segments of 256 bytes -- are brought in by bulk-async copies, double-buffered: the next 32 columns stream
        // in under the arithmetic on the current ones.  Per element the operations and their order are unchanged.
        // (Before: one row x four columns per lane with every U entry a global load inside the chain -- 1.9 of the 18.4
        // TFLOP/s of unfused FP64 on the n = 256 run; the factorisations are 95 % of that run's arithmetic.)
        const int r0 = k0 + kb;
        const int m2 = n - r0;
#ifndef DSB_COOP_OLD_UPDATE
        if (m2 > 0 && kb == DSB_COOP_NB && sc.utile != nullptr && ((uintptr_t)sc.utile & 15) == 0) {
            constexpr int NB = DSB_COOP_NB;
            double* const ubar = DSB_COOP_BAR(sc, 0);                // two barriers, one per tile buffer (the solve's: never live together)
            const int nchunk = (m2 + NB - 1) / NB;                   // column chunks of 32
            auto stage_u = [&](int ch) {                             // tile ch -> buffer ch & 1: Ut[cc * 32 + i] = U[i][c0 + cc]
                double* const Ut = sc.utile + (size_t)(ch & 1) * NB * NB;
                const int c0 = r0 + ch * NB;
                const int nc = (n - c0 < NB) ? (n - c0) : NB;
                if (bulk) {
                    if (tid == 0) {
                        coop_mbar_expect(ubar + (ch & 1), (unsigned)(nc * NB * 8));
                        for (int cc = 0; cc < nc; ++cc)
                            coop_bulk_load(Ut + (size_t)cc * NB, A + (size_t)(c0 + cc) * n + k0, (unsigned)(NB * 8), ubar + (ch & 1));
                    }
                } else {
                    for (int e = tid; e < nc * NB; e += T) Ut[e] = A[(size_t)(c0 + (e >> 5)) * n + k0 + (e & 31)];
                }
            };
            if (bulk) {
                if (tid == 0) { coop_mbar_init(ubar); coop_mbar_init(ubar + 1); }
                asm volatile("fence.proxy.async;" ::: "memory");     // step 4's writes to the U12 rows, before the async proxy reads them
                __syncthreads();
            }
            stage_u(0);
            for (int ch = 0; ch < nchunk; ++ch) {
                if (bulk) coop_mbar_wait(ubar + (ch & 1), (unsigned)((ch >> 1) & 1));
                else __syncthreads();
                if (ch + 1 < nchunk) stage_u(ch + 1);                // the other buffer: last read in chunk ch - 1 (barrier below)
                const double* const Ut = sc.utile + (size_t)(ch & 1) * NB * NB;
                const int c0 = r0 + ch * NB;
                const int ncol = (n - c0 < NB) ? (n - c0) : NB;
                // warp w takes the column groups 4 (w + nwarps j); lanes take row pairs (r, r + 32) in chunks of 64 rows.
                // The warp's (column group, row chunk) tiles form one sequence; the eight entries of A22 of the NEXT tile are
                // loaded while the current one is being updated (at 8 resident warps per SM nothing else hides that latency).
                const int nrc = (m2 + 2 * NB - 1) / (2 * NB);
                const int ncg = (ncol > 4 * wid) ? (ncol - 4 * wid + 4 * nwarps - 1) / (4 * nwarps) : 0;
                const int ntile = ncg * nrc;
                auto tile_cols = [&](int q, int& cg, int& nc4, int& ra, int& rb2) {
                    cg = 4 * wid + 4 * nwarps * (q / nrc);
                    nc4 = (ncol - cg < 4) ? (ncol - cg) : 4;
                    ra = r0 + (q % nrc) * 2 * NB + lane; rb2 = ra + NB;
                };
                double na0 = 0.0, na1 = 0.0, na2 = 0.0, na3 = 0.0, nb0 = 0.0, nb1 = 0.0, nb2 = 0.0, nb3 = 0.0;
                auto load_tile = [&](int q) {
                    int cg, nc4, ra, rb2;
                    tile_cols(q, cg, nc4, ra, rb2);
                    const bool la = ra < n, lb = rb2 < n;
                    const double* const col0 = A + (size_t)(c0 + cg + 0) * n;
                    const double* const col1 = A + (size_t)(c0 + cg + (nc4 > 1 ? 1 : 0)) * n;
                    const double* const col2 = A + (size_t)(c0 + cg + (nc4 > 2 ? 2 : 0)) * n;
                    const double* const col3 = A + (size_t)(c0 + cg + (nc4 > 3 ? 3 : 0)) * n;
                    na0 = la ? col0[ra] : 0.0; na1 = la ? col1[ra] : 0.0; na2 = la ? col2[ra] : 0.0; na3 = la ? col3[ra] : 0.0;
                    nb0 = lb ? col0[rb2] : 0.0; nb1 = lb ? col1[rb2] : 0.0; nb2 = lb ? col2[rb2] : 0.0; nb3 = lb ? col3[rb2] : 0.0;
                };
                if (ntile > 0) load_tile(0);
                for (int q = 0; q < ntile; ++q) {
                    int cg, nc4, ra, rb2;
                    tile_cols(q, cg, nc4, ra, rb2);
                    const bool la = ra < n, lb = rb2 < n;
                    double a0 = na0, a1 = na1, a2 = na2, a3 = na3, b0 = nb0, b1 = nb1, b2 = nb2, b3 = nb3;
                    if (q + 1 < ntile) load_tile(q + 1);
                    const double* const u0 = Ut + (size_t)(cg + 0) * NB;
                    const double* const u1 = Ut + (size_t)(cg + (nc4 > 1 ? 1 : 0)) * NB;
                    const double* const u2 = Ut + (size_t)(cg + (nc4 > 2 ? 2 : 0)) * NB;
                    const double* const u3 = Ut + (size_t)(cg + (nc4 > 3 ? 3 : 0)) * NB;
                    const double* const pa_ = P + (size_t)((la ? ra : r0) - k0);       // L[ra][i] at pa_[i * m]
                    const double* const pb_ = P + (size_t)((lb ? rb2 : r0) - k0);
#pragma unroll
                    for (int i = 0; i < NB; i += 2) {
                        const double2 v0 = *reinterpret_cast<const double2*>(u0 + i);
                        const double2 v1 = *reinterpret_cast<const double2*>(u1 + i);
                        const double2 v2 = *reinterpret_cast<const double2*>(u2 + i);
                        const double2 v3 = *reinterpret_cast<const double2*>(u3 + i);
                        const double la0 = pa_[(size_t)i * m], la1 = pa_[(size_t)(i + 1) * m];
                        const double lb0 = pb_[(size_t)i * m], lb1 = pb_[(size_t)(i + 1) * m];
                        a0 = (-v0.x) * la0 + a0; a1 = (-v1.x) * la0 + a1; a2 = (-v2.x) * la0 + a2; a3 = (-v3.x) * la0 + a3;
                        b0 = (-v0.x) * lb0 + b0; b1 = (-v1.x) * lb0 + b1; b2 = (-v2.x) * lb0 + b2; b3 = (-v3.x) * lb0 + b3;
                        a0 = (-v0.y) * la1 + a0; a1 = (-v1.y) * la1 + a1; a2 = (-v2.y) * la1 + a2; a3 = (-v3.y) * la1 + a3;
                        b0 = (-v0.y) * lb1 + b0; b1 = (-v1.y) * lb1 + b1; b2 = (-v2.y) * lb1 + b2; b3 = (-v3.y) * lb1 + b3;
                    }
                    double* const col0 = A + (size_t)(c0 + cg + 0) * n;
                    double* const col1 = A + (size_t)(c0 + cg + (nc4 > 1 ? 1 : 0)) * n;
                    double* const col2 = A + (size_t)(c0 + cg + (nc4 > 2 ? 2 : 0)) * n;
                    double* const col3 = A + (size_t)(c0 + cg + (nc4 > 3 ? 3 : 0)) * n;
                    if (la) { col0[ra] = a0; if (nc4 > 1) col1[ra] = a1; if (nc4 > 2) col2[ra] = a2; if (nc4 > 3) col3[ra] = a3; }
                    if (lb) { col0[rb2] = b0; if (nc4 > 1) col1[rb2] = b1; if (nc4 > 2) col2[rb2] = b2; if (nc4 > 3) col3[rb2] = b3; }
                }
                __syncthreads();                                     // tile ch is free again
            }
            if (bulk) {
                if (tid == 0) { coop_mbar_inval(ubar); coop_mbar_inval(ubar + 1); }
            }
        } else
#endif
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
        if (bulk) asm volatile("fence.proxy.async;" ::: "memory");   // the trailing columns, before the next panel's bulk copies read them
        __syncthreads();
    }
    if (bulk) {
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");              // the last write-back is in global memory
            coop_mbar_inval(fbar);
        }
        __syncthreads();
    }
    return first_bad;
}

// Solve with the factors of coop_lu_factor; b lives in SHARED memory (n doubles).  Returns false (to every
// thread) when U has a zero on its diagonal (nalgebra's solve_mut returns false; b is then unspecified).
static __device__ __noinline__ bool coop_lu_solve_ldg(const double* __restrict__ LU, int n, const int* __restrict__ piv, double* b,
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
        for (int r = rb + T; r < n; r += T) {             // further row chunks: the same load-all-then-chain pattern
#ifndef DSB_COOP_SERIAL_TAIL
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb) ? LU[(size_t)(k0 + i) * n + r] : 0.0;
            double br = b[r];
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < kb) br = (-b[k0 + i]) * lv[i] + br;
            b[r] = br;
#else
            double br = b[r];
            for (int i = 0; i < kb; ++i) br = (-b[k0 + i]) * LU[(size_t)(k0 + i) * n + r] + br;
            b[r] = br;
#endif
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
#ifndef DSB_COOP_SERIAL_TAIL
#pragma unroll
            for (int i = 0; i < 32; ++i) lv[i] = (i < kb) ? LU[(size_t)(k0 + i) * n + r] : 0.0;
            double br = b[r];
#pragma unroll
            for (int i = 31; i >= 0; --i) if (i < kb) br = (-b[k0 + i]) * lv[i] + br;
            b[r] = br;
#else
            double br = b[r];
            for (int i = kb - 1; i >= 0; --i) br = (-b[k0 + i]) * LU[(size_t)(k0 + i) * n + r] + br;
            b[r] = br;
#endif
        }
        __syncthreads();
    }
    return true;
}

// The same solve with the factors STAGED through shared memory by bulk-async copies, double-buffered: panels of 16
// columns (two of them fill the n x 32 panel scratch); while the block works on panel k out of one buffer, the 16 column
// segments of panel k + 1 stream into the other and complete on its mbarrier.  Forward pass: panel = columns
// [k0, k0 + 16), rows [k0, n); backward pass: the same columns, rows [0, k0 + 16).  Arithmetic and its order per element
// are those of coop_lu_solve_ldg (blocking only regroups the column-axpy steps).  Needs the panel scratch (sc.panel,
// n x 32 doubles) and n even; otherwise the plain-load version runs.
static __device__ __noinline__ bool coop_lu_solve(const double* __restrict__ LU, int n, const int* __restrict__ piv, double* b,
                              const CoopScratch& sc, bool have_panel = true) {
#if defined(DSB_COOP_NO_BULK) || !defined(DSB_COOP_BULK_SOLVE)
    // Measured on one B200 (heat-equation DAE n = 256 through the dense path, 4096 instances): 1467 ms with this staged solve,
    // 1217 ms with the plain-load one -- the two triangular solves are 5 % of the arithmetic of that run (the factorisations
    // are the rest), and panels of 16 double the number of diagonal-block chains and barriers.  Kept behind
    // DSB_COOP_BULK_SOLVE; the bulk copies earn their keep in the factorisation (panel staging, write-back, U tiles).
    (void)have_panel;
    return coop_lu_solve_ldg(LU, n, piv, b, sc);
#else
    if (!have_panel || (n & 1) || ((uintptr_t)LU & 15) != 0) return coop_lu_solve_ldg(LU, n, piv, b, sc);
    constexpr int NBS = 16;
    const int tid = threadIdx.x, T = blockDim.x;
    double* const bar0 = DSB_COOP_BAR(sc, 0);
    double* const buf0 = sc.panel;
    double* const buf1 = sc.panel + (size_t)n * NBS;
    const int nblk = (n + NBS - 1) / NBS;
    // iteration it: forward panel it (it < nblk), then backward panel 2 nblk - 1 - it
    auto issue = [&](int it) {                          // thread 0 only
        const bool fwd = it < nblk;
        const int kbk = fwd ? it : 2 * nblk - 1 - it;
        const int k0 = kbk * NBS;
        const int kb = (n - k0 < NBS) ? (n - k0) : NBS;
        const int r_lo = fwd ? k0 : 0;
        const int m = fwd ? (n - k0) : (k0 + kb);
        double* const P = (it & 1) ? buf1 : buf0;
        double* const bar = bar0 + (it & 1);
        coop_mbar_expect(bar, (unsigned)(kb * m * 8));
        for (int c = 0; c < kb; ++c) coop_bulk_load(P + (size_t)c * m, LU + (size_t)(k0 + c) * n + r_lo, (unsigned)(m * 8), bar);
    };
    if (tid == 0) {
        for (int i = 0; i < n; ++i) { const int p = piv[i]; if (p != i) { const double tmp = b[i]; b[i] = b[p]; b[p] = tmp; } }
        sc.bcast[2] = 1;
        coop_mbar_init(bar0); coop_mbar_init(bar0 + 1);
        issue(0);
    }
    __syncthreads();
    bool ok_all = true;
    const int nit = 2 * nblk;
    for (int it = 0; it < nit; ++it) {
        if (tid == 0 && it + 1 < nit) issue(it + 1);      // the other buffer: last read in iteration it - 1, which ended with a barrier
        const bool fwd = it < nblk;
        const int kbk = fwd ? it : 2 * nblk - 1 - it;
        const int k0 = kbk * NBS;
        const int kb = (n - k0 < NBS) ? (n - k0) : NBS;
        const int m = fwd ? (n - k0) : (k0 + kb);
        const double* const P = (it & 1) ? buf1 : buf0;
        coop_mbar_wait(bar0 + (it & 1), (unsigned)((it >> 1) & 1));
        if (fwd) {
            // diagonal block: unit lower triangle, lane = row within the block; P[i * m + lr], lr = row - k0
            if (tid < 32) {
                const int lr = tid;
                double lv[NBS];
#pragma unroll
                for (int i = 0; i < NBS; ++i) lv[i] = (i < kb && lr < kb) ? P[(size_t)i * m + lr] : 0.0;
                double br = (lr < kb) ? b[k0 + lr] : 0.0;
#pragma unroll
                for (int i = 0; i < NBS; ++i) {
                    const double coeff = __shfl_sync(0xffffffffu, br, i);
                    if (i < kb && lr > i && lr < kb) br = (-coeff) * lv[i] + br;
                }
                if (lr < kb) b[k0 + lr] = br;
            }
            __syncthreads();
            for (int r = k0 + kb + tid; r < n; r += T) {   // rows below: contributions in ASCENDING i
                double br = b[r];
#pragma unroll
                for (int i = 0; i < NBS; ++i) if (i < kb) br = (-b[k0 + i]) * P[(size_t)i * m + (r - k0)] + br;
                b[r] = br;
            }
        } else {
            // diagonal block: upper triangle with its diagonal, from the bottom; P[i * m + r], r = absolute row
            if (tid < 32) {
                const int lr = tid;
                double lv[NBS];
#pragma unroll
                for (int i = 0; i < NBS; ++i) lv[i] = (i < kb && lr < kb) ? P[(size_t)i * m + (k0 + lr)] : 0.0;
                double br = (lr < kb) ? b[k0 + lr] : 0.0;
                bool ok = true;
#pragma unroll
                for (int i = NBS - 1; i >= 0; --i) {
                    if (i < kb && ok) {
                        const double diag = __shfl_sync(0xffffffffu, lv[i], i);       // U[i][i] lives in lane i
                        if (diag == 0.0) ok = false;                                    // uniform across the warp
                        else {
                            double coeff = 0.0;
                            if (lr == i) { coeff = br / diag; br = coeff; }
                            coeff = __shfl_sync(0xffffffffu, coeff, i);
                            if (lr < i) br = (-coeff) * lv[i] + br;
                        }
                    }
                }
                if (lr < kb) b[k0 + lr] = br;
                if (!ok && tid == 0) sc.bcast[2] = 0;
            }
            __syncthreads();
            if (sc.bcast[2] == 0) { ok_all = false; }
            else {
                for (int r = tid; r < k0; r += T) {        // rows above: contributions in DESCENDING i
                    double br = b[r];
#pragma unroll
                    for (int i = NBS - 1; i >= 0; --i) if (i < kb) br = (-b[k0 + i]) * P[(size_t)i * m + r] + br;
                    b[r] = br;
                }
            }
        }
        __syncthreads();
        if (!ok_all) {
            // a zero on U's diagonal: drain the copy in flight, then give up (b is unspecified, as in nalgebra)
            if (it + 1 < nit) coop_mbar_wait(bar0 + ((it + 1) & 1), (unsigned)(((it + 1) >> 1) & 1));
            break;
        }
    }
    __syncthreads();
    if (tid == 0) { coop_mbar_inval(bar0); coop_mbar_inval(bar0 + 1); }
    __syncthreads();
    return ok_all;
#endif
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

