// dsb_wband_bdf_kernel.cuh -- `problem.bdf::<LS>()?.solve_dense(t_eval)` for BANDED systems of medium size (n > 16,
// df/dy and M inside a declared band kl, ku <= 2), ONE WARP PER INSTANCE with the instance's whole working set in
// SHARED memory.
//
// Why (round-1 profiles of the one-lane-per-instance kernel dsb_band_bdf_kernel.cuh, profiles/r1_s4_band_*): with the
// state in global memory every component of every vector operation was a global round trip (long-scoreboard stalls of
// 9-24 cycles per issue at 6-24 % resident warps, 1.4-2.5x the algorithmic DRAM traffic).  Here
//   * the Newton work vectors, the band factors, their pivots and RN(1 / U_jj) live in the warp's shared-memory slice
//     for the whole integration of the instance (10.5 n words for a tridiagonal system); HBM sees the parameters, the
//     initial state and the result columns, nothing else;
//   * the difference array D (n x 8) joins them in shared memory where that still leaves 8 instances per SM (n = 42);
//     for larger systems it lives in a per-warp global-memory slot that stays in L2 (the resident warps' slots are a few
//     tens of MB), is read in chunks with L1 bypassed, and costs one L2 round trip per chunk in the ~5 passes a step
//     makes over it -- which buys 10 instead of 6 instances per SM at n = 256 (measured 1.55x);
//   * df/dy (and M for DAEs) are only needed when the iteration matrix is rebuilt: they are PARKED in the same slot (3 n
//     words written per Jacobian evaluation) and df/dy comes back by one bulk-async copy (cp.async.bulk + mbarrier, the
//     1-D TMA path) straight into the band rows where A = M - c J is assembled;
//   * every vector operation is lane-parallel (component i on lane i mod 32, stride-1 shared-memory rows: conflict-free);
//   * the inherently sequential recurrences -- band LU, forward / back substitution, the in-order sum of a weighted
//     norm (the reference adds its terms sequentially, vector/nalgebra_serial.rs:395-408) -- run on one lane out of
//     shared memory with the running entries in a register window.  Their cost is a pure FP64 dependency chain (8.2
//     cycles per DADD / DMUL / DFMA on B200, an IEEE division 110: tools/fp64_peak.cu), so the back substitution divides
//     through the stored reciprocal -- 3 dependent operations, proven correct behind the chain (SmemBandLU::solve) --, its
//     rows are branch-free, and the forward sweep walks interchange-free segments between the few rows that pivoted;
//   * result columns are staged in a dead work vector and leave through a bulk-async store (cp.async.bulk
//     shared -> global) that drains while the warp integrates on; outputs are written INSTANCE-major (the reference's
//     host layout), which makes every column one contiguous run.  Equations with an output function and declared
//     dependencies (the battery model's terminal voltage) evaluate up to 32 pending columns at once, one per lane.
// Control flow is uniform over the warp, so the per-lane state machine of dsb_band_bdf_kernel.cuh is kept block by block
// (same restated functions, same expression order, bit-identical results: tests/test_gpu_warp_band_parity.py) without
// its warp scheduler.  Warps are persistent and draw instances from a global work counter.
// What bounds it (profiles/r2_wband_*): the dependency chains of the one-lane sections at 8-12 resident warps per SM
// (issue slots 33 % busy, `wait` = fixed-latency dependency is the dominant stall); the number of resident warps is set by
// shared memory (n >= 200) or registers.
//
// Restated functions: the list of dsb_bdf_kernel.cuh plus new_without_initialise / set_step_size
// (ode_solver/state.rs:1086-1124, 1209-1277).  Not built here (the launcher falls back to the one-lane-per-instance
// kernel): reset functions, (E)SDIRK.
#pragma once
#include "dsb_band_bdf_kernel.cuh"

#if defined(__CUDACC__)
#define DSB_WLANES 32
#else
#define DSB_WLANES 1                // tests/host_emu: the same source, one lane per warp
#endif
// Warps per block.  Registers are allotted per scheduler (4 x 16384): 16 warps leave 128 registers per lane, 9-12 leave
// 168, 8 or fewer 255.
#ifndef DSB_WBAND_MAX_WARPS
#define DSB_WBAND_MAX_WARPS 16              // small systems, D in shared memory
#endif
#ifndef DSB_WBAND_MAX_WARPS_D_GLOBAL
#define DSB_WBAND_MAX_WARPS_D_GLOBAL 12     // larger systems, D in the global-memory slot
#endif
#ifndef DSB_WBAND_MIN_WARPS_D_SHARED
#define DSB_WBAND_MIN_WARPS_D_SHARED 8
#endif
#ifndef DSB_WBAND_CHUNK
#define DSB_WBAND_CHUNK 2           // components per lane whose difference-array loads go out together (dsb_warp_for)
#endif
#define DSB_WBAND_SMEM_BYTES (227 * 1024)

template <class M, class = void> struct dsb_wband_max_warps { static constexpr int value = 64; };
template <class M> struct dsb_wband_max_warps<M, decltype((void)M::WBAND_MAX_WARPS)> { static constexpr int value = M::WBAND_MAX_WARPS; };

template <class M, class = void> struct dsb_wband_chunk { static constexpr int value = DSB_WBAND_CHUNK; };
template <class M> struct dsb_wband_chunk<M, decltype((void)M::WBAND_CHUNK)> { static constexpr int value = M::WBAND_CHUNK; };

template <class M>
struct WBandLayout {
    static constexpr int N = M::N, NP = M::NP;
    static constexpr int KL = M::BAND_KL, KU = M::BAND_KU, KV = KL + KU;
    static constexpr int LDJ = KL + KU + 1;                         // band rows of df/dy (and of M)
    static constexpr int LDAB = 2 * KL + KU + 1;                    // band rows of the factors (kl rows of fill-in)
    static constexpr int NS = (N + 1) & ~1;                         // row stride: even, so every row is 16-byte aligned
    // shared-memory words of one warp; a band row r holds one diagonal: entry (i, j) of row KV + i - j (factors) or
    // KU + i - j (df/dy, M) sits at column j
    // The difference array D (n x 8) takes 8 of the 18.5 n words of an instance.  Where the whole set still lets
    // DSB_WBAND_MIN_WARPS_D_SHARED instances share an SM (n = 42: 16), D stays in shared memory; for larger systems it
    // moves to the warp's global-memory slot (L2-resident, loaded in chunks, L1 bypassed), which nearly doubles the
    // instances per SM (n = 256: 6 -> 10; n = 200: 7 -> 12) at the price of one L2 round trip per chunk in the five or so
    // passes over D that a step makes.
    static constexpr int WORDS_NO_D = 5 * NS + LDAB * NS + NS + NS / 2 + 26 + 1 + 5;
    static constexpr int BLOCK_WORDS = (NS + 15) / 16 * 16;         // shared by the block's warps: the absolute tolerances
    static constexpr int FIT_D_SHARED = (DSB_WBAND_SMEM_BYTES - BLOCK_WORDS * 8) / (((WORDS_NO_D + DSB_NDIFF * NS + 15) / 16 * 16) * 8);
    static constexpr bool D_SHARED = FIT_D_SHARED >= DSB_WBAND_MIN_WARPS_D_SHARED;
    static constexpr int O_D = 0;                                   // D[DSB_NDIFF][NS] (when D_SHARED)
    static constexpr int O_Y = D_SHARED ? DSB_NDIFF * NS : 0;       // state.y
    // the predictor and psi - y_predict are only ever read component by component: for the larger systems they follow D
    // into the global-memory slot (two more L2 passes per Newton iteration, 12 instead of 10 instances per SM at n = 256)
    static constexpr int O_YC = O_Y + NS;                           // Newton iterate
    static constexpr int O_DL = O_YC + NS;                          // Newton residual / update, norm terms, output staging
    // ... unless the warps are capped below what shared memory would hold anyway (M::WBAND_MAX_WARPS: the battery model)
    static constexpr int MAXW_MODEL = dsb_wband_max_warps<M>::value;
    static constexpr int MAXW_KIND = D_SHARED ? DSB_WBAND_MAX_WARPS : DSB_WBAND_MAX_WARPS_D_GLOBAL;
    static constexpr int MAXW = MAXW_MODEL < MAXW_KIND ? MAXW_MODEL : MAXW_KIND;
    static constexpr int FIT_VEC_SHARED = (DSB_WBAND_SMEM_BYTES - BLOCK_WORDS * 8) / (((WORDS_NO_D + 15) / 16 * 16) * 8);
    static constexpr bool VEC_SHARED = D_SHARED || FIT_VEC_SHARED >= MAXW;
    static constexpr int O_YP = O_DL + NS;                          // y_predict (when VEC_SHARED)
    static constexpr int O_PSI = O_YP + NS;                         // psi - y_predict (when VEC_SHARED)
    static constexpr int O_VEC_END = VEC_SHARED ? O_PSI + NS : O_DL + NS;
    static constexpr int O_AB = O_VEC_END;                          // factors [LDAB][NS]
    static constexpr int O_RCP = O_AB + LDAB * NS;                  // RN(1 / U_jj) (dsb_math.h: dsb_rcp)
    static constexpr int O_PIV = O_RCP + NS;                        // int32 pivot offsets (row j interchanged with row j + piv[j])
    static constexpr int O_RU = O_PIV + NS / 2;                     // rows / columns 1..5 of R U (rescale)
    static constexpr int O_BAR = O_RU + 26;                         // mbarrier of the bulk-async loads
    static constexpr int O_FLAG = O_BAR + 1;                        // int32: zero pivots met by the last factorisation, its row
                                                                    // interchanges, the rows of the first 8 of them
    static constexpr int WORDS = (O_FLAG + 5 + 15) / 16 * 16;       // slices start on 128-byte boundaries
    static constexpr int FIT = (DSB_WBAND_SMEM_BYTES - BLOCK_WORDS * 8) / (WORDS * 8);
    // (MAXW and the chunk size of the passes over D can be set per equation set, M::WBAND_MAX_WARPS / M::WBAND_CHUNK: what
    // matters is that the kernel stays inside the registers the warp count leaves -- 168 per lane at 9-12 warps -- because
    // a spilled kernel's local-memory footprint does not fit the few KB of L1 next to 200+ KB of shared memory; see the
    // battery model's figures in dsb_models.h)
    static constexpr int WARPS = FIT < 1 ? 1 : (FIT < MAXW ? FIT : MAXW);
    static constexpr int THREADS = WARPS * DSB_WLANES;
    static constexpr size_t SMEM_BYTES = (size_t)(BLOCK_WORDS + WORDS * WARPS) * 8;
    static constexpr bool FITS = FIT >= 1;
    // global-memory slot of one warp (L2-resident: the resident warps' slots are a few tens of MB): the difference array
    // D[DSB_NDIFF][NS] (unless D_SHARED), df/dy band [LDJ][NS], then M band [LDJ][NS] (DAEs)
    static constexpr int G_D = 0;
    static constexpr int G_YP = G_D + (D_SHARED ? 0 : DSB_NDIFF * NS);
    static constexpr int G_PSI = G_YP + (VEC_SHARED ? 0 : NS);
    static constexpr int G_J = G_PSI + (VEC_SHARED ? 0 : NS);
    static constexpr int G_M = G_J + LDJ * NS;
    static constexpr int G_WORDS = G_M + (M::HAS_MASS ? LDJ * NS : 0);
    static_assert(KL >= 1 && KL <= 2 && KU >= 1 && KU <= 2, "register windows are sized for kl, ku <= 2");
};

// ---- warp plumbing (trivial with one lane on the host) ------------------------------------------------------------------
DSB_DEV int dsb_wlane() {
#if defined(__CUDA_ARCH__)
    return (int)(threadIdx.x & 31u);
#else
    return 0;
#endif
}
DSB_DEV void dsb_wsync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
template <class T> DSB_DEV T dsb_wbcast(T v, int src) {
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xffffffffu, v, src);
#else
    (void)src; return v;
#endif
}
DSB_DEV unsigned dsb_wballot(bool p) {
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xffffffffu, p);
#else
    return p ? 1u : 0u;
#endif
}

// ---- bulk-async copies (1-D TMA): issued and waited for by ONE lane ------------------------------------------------------
// store: shared -> global, completion tracked by the issuing thread's bulk groups; load: global -> shared, completion
// on an mbarrier in shared memory.  Sizes and addresses are multiples of 16 bytes (NS is even).
DSB_DEV void dsb_bulk_store(double* gdst, const double* ssrc, int words) {
#if defined(__CUDA_ARCH__)
    const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the lanes' generic-proxy writes, ordered by __syncwarp
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(words * 8) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
    for (int i = 0; i < words; ++i) gdst[i] = ssrc[i];
#endif
}
template <int PENDING> DSB_DEV void dsb_bulk_store_wait_read() {   // all but PENDING groups have finished READING shared memory
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
#endif
}
DSB_DEV void dsb_bulk_store_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
DSB_DEV void dsb_mbar_init(double* bar) {
#if defined(__CUDA_ARCH__)
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    (void)bar;
#endif
}
// one lane: expect `words` doubles, then copy them global -> shared
DSB_DEV void dsb_bulk_load(double* sdst, const double* gsrc, int words, double* bar) {
#if defined(__CUDA_ARCH__)
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), s = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("fence.proxy.async;" ::: "memory");                 // earlier generic-proxy accesses to both ends
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(words * 8) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s), "l"(gsrc), "r"(words * 8), "r"(b) : "memory");
#else
    (void)bar;
    for (int i = 0; i < words; ++i) sdst[i] = gsrc[i];
#endif
}
// every lane: wait for the phase with the given parity
DSB_DEV void dsb_mbar_wait(double* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
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
#else
    (void)bar; (void)parity;
#endif
}

// The difference array, df/dy and M live in the warp's global-memory slot and stay in L2: their loads and stores bypass
// L1 (ld.global.cg / st.global.cg), which -- next to 200+ KB of shared memory -- is only a few tens of KB and holds the
// equations' coefficient tables and the column metadata (streaming D through it evicted them: 1.45x slower on n = 200).
DSB_DEV double dsb_ld_l2(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
DSB_DEV void dsb_st_l2(double* p, double v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}

// indexable view of a shared-memory vector (what the component-wise equations read)
struct WVec {
    const double* base;
    __device__ __forceinline__ double operator[](int k) const { return base[k]; }
};
// y + (psi - y0), the argument of the mass matrix in the BDF residual (op/bdf.rs:240-256), formed on the fly
template <bool B_GLOBAL>
struct WSumVec {
    const double* a; const double* b;            // a in shared memory; b in shared memory or (B_GLOBAL) in the warp's global slot
    __device__ __forceinline__ double operator[](int k) const { return a[k] + (B_GLOBAL ? dsb_ld_l2(b + k) : b[k]); }
};
// the NDEP state components an output / root function declares, held in registers
template <class M, int NDEP>
struct WDepVec {
    double v[NDEP > 0 ? NDEP : 1];
    __device__ __forceinline__ double operator[](int k) const {
        double r = v[0];
#pragma unroll
        for (int q = 1; q < NDEP; ++q) r = (k == M::dep(q)) ? v[q] : r;
        return r;
    }
};

// Lane-strided component loop in chunks: load(i) for U components of the lane first, then use(i, value) for each.  The
// difference array lives in the warp's global-memory slot (L2): with the loads of a chunk issued back to back a pass over
// it costs about one L2 round trip per chunk instead of one per component.  load() must not read what use() writes for
// another component.
template <int U, class R, class L, class F>
DSB_DEV void dsb_warp_for(const int lane, const int n, L&& load, F&& use) {
#pragma unroll 1
    for (int i0 = lane; i0 < n; i0 += U * DSB_WLANES) {
        R r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * DSB_WLANES; if (i < n) r[u] = load(i); }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * DSB_WLANES; if (i < n) use(i, r[u]); }
    }
}
struct WCols { double v[DSB_MAX_ORDER + 2]; };
struct WColsYp { WCols c; double yp; };

// ---- band LU on a shared-memory band, ONE lane ----------------------------------------------------------------------------
// The arithmetic of dsb_band_lu.cuh (= nalgebra 0.35 `DMatrix::lu()` / `LU::solve_mut`, only operations with an exactly
// zero operand skipped) on the row-per-diagonal layout: entry (i, j) at ab[(KV + i - j) * NS + j].
template <int N, int NS, int KL, int KU, int U = 4>
struct SmemBandLU {
    static constexpr int KV = KL + KU, LDAB = 2 * KL + KU + 1;
#define AB_(j, r) ab[(r) * NS + (j)]
    // returns the number of exactly zero pivots (the factors are then unusable: LaError::LuSolveFailed at the next solve);
    // *nswaps = number of row interchanges, swaprow[0 .. min(nswaps, 8)) their rows (see solve(): FWD)
    static DSB_DEV int factor(double* __restrict__ ab, int* __restrict__ piv, double* __restrict__ rcp, int* __restrict__ nswaps,
                              int* __restrict__ swaprow) {
        int nzero = 0, nsw = 0;
        int jlast = 0;                                   // last column touched by the fill-in so far
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const int km = (KL < N - 1 - j) ? KL : (N - 1 - j);
            double colv[KL + 1];
#pragma unroll
            for (int d = 0; d <= KL; ++d) colv[d] = (d <= km) ? AB_(j, KV + d) : 0.0;
            int jp = 0;
            double best = -1.0;
#pragma unroll
            for (int d = 0; d <= KL; ++d) {
                const double av = dsb_abs(colv[d]);
                if (d <= km && av == av && av > best) { best = av; jp = d; }     // first maximum, NaNs never win
            }
            if (colv[0] != colv[0]) jp = 0;                  // a NaN diagonal keeps the diagonal
            double diag = colv[0];
#pragma unroll
            for (int d = 1; d <= KL; ++d) if (jp == d) diag = colv[d];
            if (diag == 0.0) { piv[j] = 0; rcp[j] = 0.0; ++nzero; continue; }
            piv[j] = jp;
            { const int cand = (j + KU + jp < N - 1) ? (j + KU + jp) : (N - 1); if (cand > jlast) jlast = cand; }
            if (jp != 0) {
                if (nsw < 8) swaprow[nsw] = j;               // the first MAXSW interchange rows, for the forward sweep
                ++nsw;
#pragma unroll
                for (int q = 0; q <= KV; ++q) {              // columns j .. jlast (at most kv + 1 of them)
                    const int cq = j + q;
                    if (cq <= jlast) {
                        const double a = AB_(cq, KV - q), b = AB_(cq, KV - q + jp);
                        AB_(cq, KV - q) = b; AB_(cq, KV - q + jp) = a;
                    }
                }
#pragma unroll
                for (int d = 0; d <= KL; ++d) {              // the register copy of column j follows the interchange
                    const double a = colv[0];
                    if (jp == d && d != 0) { colv[0] = colv[d]; colv[d] = a; }
                }
            }
            const double inv_diag = 1.0 / colv[0];
            rcp[j] = dsb_rcp_from_narrow(colv[0], inv_diag);
            if (km > 0) {
#pragma unroll
                for (int d = 1; d <= KL; ++d) if (d <= km) { colv[d] *= inv_diag; AB_(j, KV + d) = colv[d]; }
#pragma unroll
                for (int q = 1; q <= KV; ++q) {              // columns j + 1 .. jlast
                    const int cq = j + q;
                    if (cq <= jlast) {
                        const double mpk = -AB_(cq, KV - q);
#pragma unroll
                        for (int d = 1; d <= KL; ++d)
                            if (d <= km) AB_(cq, KV - q + d) = mpk * colv[d] + AB_(cq, KV - q + d);
                    }
                }
            }
        }
        *nswaps = nsw;
        return nzero;
    }

    // b <- A^-1 b.  Returns 1, or 2 when the FAST back substitution cannot vouch for one of its quotients: the caller then
    // rebuilds the right-hand side and calls the EXACT form (plain IEEE divisions), which always returns 1.  (A zero pivot
    // -- LaError::LuSolveFailed -- is recorded by factor(); the caller does not call solve() then.)
    //
    // Both sweeps are pure dependency chains on one lane, so what counts is the latency per row:
    //   * blocks of U rows: the block's pivots, multipliers and incoming right-hand-side entries are loaded first, the
    //     recurrence runs in registers, the block's results are stored last -- shared-memory latency stays off the chain;
    //     the rows that need bounds tests (the last kl of the forward sweep, the last kl + ku of the backward one) run in
    //     a separate tail loop;
    //   * FAST back substitution: x_i = w / U_ii continues with q1 = RN(q0 + RN(w - q0 U_ii) r), q0 = RN(w r), r the
    //     stored RN(1 / U_ii) (3 dependent operations).  q1 is within half an ulp + 2^-104 of the quotient, i.e. it IS
    //     RN(w / U_ii) unless the quotient lies that close to a rounding boundary; one more residual step
    //     q2 = RN(q1 + RN(w - q1 U_ii) r) is the correctly rounded quotient (dsb_math.h: dsb_div_rcp), so q2 == q1 proves
    //     q1 right.  That test runs BEHIND the chain (nothing waits for it); a refuted quotient, a NaN reciprocal
    //     (dsb_rcp_from_narrow: |exponent of U_ii| > 100) or a non-zero numerator outside the proof's range (dsb_math.h:
    //     dsb_numerator_in_wide_range -- subnormals and the last 30 decades above them, infinities, NaNs) only raise a
    //     flag that is looked at once per solve.
    // FWD: how the forward sweep meets row interchanges -- 0: the factorisation made none; 1: a few (at most MAXSW, their
    // rows listed in ascending order in swaprow[0 .. nsw)): interchange-free segments between them, so no select sits on
    // the chain; 2: any number, tested row by row (selects).
    static constexpr int MAXSW = 8;
    template <bool EXACT, int FWD>
    static DSB_DEV int solve(const double* __restrict__ ab, const int* __restrict__ piv, const double* __restrict__ rcp,
                             double* __restrict__ b, const int* __restrict__ swaprow, const int nsw) {
        {
            double w[KL + 1];
#pragma unroll
            for (int d = 0; d <= KL; ++d) w[d] = b[d];
            // one row of the forward sweep: interchange, b[j] leaves the window, the kl multipliers of column j act
            auto fwd_row = [&](auto SW, const int jp, const double (&lm_)[KL], const double bin) -> double {
                if constexpr (decltype(SW)::value) {
                    if (jp != 0) {
                        const double a = w[0];
#pragma unroll
                        for (int d = 1; d <= KL; ++d) if (jp == d) { w[0] = w[d]; w[d] = a; }
                    }
                }
                const double bj = w[0];
                const double nbj = -bj;
#pragma unroll
                for (int d = 1; d <= KL; ++d) w[d] = nbj * lm_[d - 1] + w[d];
#pragma unroll
                for (int d = 0; d < KL; ++d) w[d] = w[d + 1];
                w[KL] = bin;
                return bj;
            };
            // one row with its bounds tests (multipliers below the matrix are exact zeros)
            auto fwd_single = [&](auto SW, const int j) {
                double lm_[KL];
#pragma unroll
                for (int d = 1; d <= KL; ++d) lm_[d - 1] = (j + d < N) ? AB_(j, KV + d) : 0.0;
                const double bin = (j + 1 + KL < N) ? b[j + 1 + KL] : 0.0;
                b[j] = fwd_row(SW, decltype(SW)::value ? piv[j] : 0, lm_, bin);
            };
            // rows [j0, j1), j1 <= N - 1: blocks of U rows whose multipliers and incoming entry all exist, then single rows
            auto fwd_range = [&](auto SW, int j0, const int j1) {
                const int safe = (j1 < N - 1 - KL) ? j1 : (N - 1 - KL);
#pragma unroll 1
                for (; j0 + U <= safe; j0 += U) {
                    int jp[U];
                    double lm_[U][KL], bin[U], out[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        jp[u] = decltype(SW)::value ? piv[j0 + u] : 0;
#pragma unroll
                        for (int d = 1; d <= KL; ++d) lm_[u][d - 1] = AB_(j0 + u, KV + d);
                        bin[u] = b[j0 + u + 1 + KL];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) out[u] = fwd_row(SW, jp[u], lm_[u], bin[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u) b[j0 + u] = out[u];
                }
#pragma unroll 1
                for (; j0 < j1; ++j0) fwd_single(SW, j0);
            };
            if constexpr (FWD == 0) {
                fwd_range(std::false_type{}, 0, N - 1);
            } else if constexpr (FWD == 1) {
                int j = 0;
#pragma unroll 1
                for (int k = 0; k < nsw; ++k) {
                    const int r = swaprow[k];
                    fwd_range(std::false_type{}, j, r);
                    fwd_single(std::true_type{}, r);
                    j = r + 1;
                }
                fwd_range(std::false_type{}, j, N - 1);
            } else {
                fwd_range(std::true_type{}, 0, N - 1);
            }
            b[N - 1] = w[0];
        }
        int bad = 0;
        {
            double w[KV + 1];
#pragma unroll
            for (int e = 0; e <= KV; ++e) w[e] = b[N - 1 - e];
            // one row of the backward sweep: x_i = w[0] / U_ii, then column i of U acts on the kv entries above
            // The fast form is branch-free: the sign of a zero quotient is put in by a bit operation (sign(a) ^ sign(U_ii) is
            // the sign of every correctly rounded quotient), and what cannot be vouched for -- q2 != q1, a numerator outside
            // the proof's range -- only accumulates into `bad`, behind the chain.
            auto bwd_row = [&](const double (&up)[KV + 1], const double rc, const double bin) -> double {
                const double a = w[0], diag = up[0];
                double x;
                if constexpr (EXACT) {
                    x = a / diag;
                } else {
                    const double q0 = a * rc;
                    const double e0 = dsb_fma(-q0, diag, a);
                    const double q1 = dsb_fma(e0, rc, q0);
                    x = dsb_quotient_sign(q1, a, diag);
                    const double e1 = dsb_fma(-q1, diag, a);
                    const double q2 = dsb_fma(e1, rc, q1);
                    bad |= (int)!(q2 == q1) | ((int)(a != 0.0) & (int)!dsb_numerator_in_wide_range(a));    // no short circuits: no branches
                }
                const double nx = -x;
#pragma unroll
                for (int e = 1; e <= KV; ++e) w[e] = nx * up[e] + w[e];
#pragma unroll
                for (int e = 0; e < KV; ++e) w[e] = w[e + 1];
                w[KV] = bin;
                return x;
            };
            int i0 = N - 1;
#pragma unroll 1
            for (; i0 - (U - 1) >= KV + 1; i0 -= U) {           // rows whose kv upper entries and incoming entry all exist
                double up[U][KV + 1], rc[U], bin[U], out[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 - u;
#pragma unroll
                    for (int e = 0; e <= KV; ++e) up[u][e] = AB_(i, KV - e);
                    rc[u] = rcp[i];
                    bin[u] = b[i - 1 - KV];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) out[u] = bwd_row(up[u], rc[u], bin[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) b[i0 - u] = out[u];
            }
#pragma unroll 1
            for (int i = i0; i >= 0; --i) {                     // tail: entries above the matrix are exact zeros
                double up[KV + 1];
#pragma unroll
                for (int e = 0; e <= KV; ++e) up[e] = (i - e >= 0) ? AB_(i, KV - e) : 0.0;
                const double bin = (i - 1 - KV >= 0) ? b[i - 1 - KV] : 0.0;
                b[i] = bwd_row(up, rcp[i], bin);
            }
        }
        return bad ? 2 : 1;
    }
#undef AB_
};

template <class M>
__global__ void __launch_bounds__(WBandLayout<M>::THREADS, 1)
dsb_wband_bdf_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa, const __grid_constant__ DsbBatchBuffers bb,
                                 const __grid_constant__ DsbBandMeta meta, double* __restrict__ ws, double* __restrict__ ys_im,
                                 unsigned long long* __restrict__ work_counter) {
    typedef WBandLayout<M> Lay;
    constexpr int N = Lay::N, NP = Lay::NP, NS = Lay::NS, KL = Lay::KL, KU = Lay::KU, KV = Lay::KV, LDJ = Lay::LDJ, LDAB = Lay::LDAB;
    constexpr int LANES = DSB_WLANES;
    typedef SmemBandLU<N, NS, KL, KU> BLU;
    extern __shared__ double dsb_lane_smem[];
    const int lane = dsb_wlane();
    const int warp = (int)(threadIdx.x / LANES);
    double* const satol = dsb_lane_smem;                             // [N], one copy per block
    double* const sm = dsb_lane_smem + Lay::BLOCK_WORDS + (size_t)warp * Lay::WORDS;
    double* const gslot = ws + ((size_t)blockIdx.x * Lay::WARPS + warp) * Lay::G_WORDS;
#define SD_LD(j, i) (Lay::D_SHARED ? sm[Lay::O_D + (j) * NS + (i)] : dsb_ld_l2(&gslot[Lay::G_D + (j) * NS + (i)]))
#define SD_ST(j, i, v) do { if (Lay::D_SHARED) sm[Lay::O_D + (j) * NS + (i)] = (v); else dsb_st_l2(&gslot[Lay::G_D + (j) * NS + (i)], (v)); } while (0)
#define SY(i) sm[Lay::O_Y + (i)]
#define SYP_LD(i) (Lay::VEC_SHARED ? sm[Lay::O_YP + (i)] : dsb_ld_l2(&gslot[Lay::G_YP + (i)]))
#define SYP_ST(i, v) do { if (Lay::VEC_SHARED) sm[Lay::O_YP + (i)] = (v); else dsb_st_l2(&gslot[Lay::G_YP + (i)], (v)); } while (0)
#define SYC(i) sm[Lay::O_YC + (i)]
#define SPSI_LD(i) (Lay::VEC_SHARED ? sm[Lay::O_PSI + (i)] : dsb_ld_l2(&gslot[Lay::G_PSI + (i)]))
#define SPSI_ST(i, v) do { if (Lay::VEC_SHARED) sm[Lay::O_PSI + (i)] = (v); else dsb_st_l2(&gslot[Lay::G_PSI + (i)], (v)); } while (0)
#define SDL(i) sm[Lay::O_DL + (i)]
#define SAB(j, r) sm[Lay::O_AB + (r) * NS + (j)]
#define SRU(i, j) sm[Lay::O_RU + ((i) - 1) * 5 + ((j) - 1)]
#define GJ(j, r) gslot[Lay::G_J + (r) * NS + (j)]
#define GM(j, r) gslot[Lay::G_M + (r) * NS + (j)]
#define DSB_DIV(a, b) DsbDivShared::div((a), (b))
#define WFOR(i) for (int i = lane; i < N; i += LANES)
    int* const spiv = reinterpret_cast<int*>(sm + Lay::O_PIV);
    int* const sflag = reinterpret_cast<int*>(sm + Lay::O_FLAG);
    double* const sbar = sm + Lay::O_BAR;
    const WVec vY{sm + Lay::O_Y}, vYC{sm + Lay::O_YC}, vDL{sm + Lay::O_DL};
    const WSumVec<!Lay::VEC_SHARED> vTMP{sm + Lay::O_YC, Lay::VEC_SHARED ? sm + Lay::O_PSI : gslot + Lay::G_PSI};

    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const double eps = 2.220446049250313e-16;
    constexpr int NOUT = dsb_model_nout<M>::value;
    constexpr int NDEP = dsb_model_ndep<M>::value;
    constexpr int NR = dsb_model_nroots<M>::value;
    constexpr int DCHUNK = dsb_wband_chunk<M>::value;                      // passes over the difference array (global loads unless D_SHARED)
    constexpr int VCHUNK = Lay::VEC_SHARED ? 1 : 2 * DCHUNK;     // predictor / psi passes: chunked only when they are global loads
    constexpr bool BULK_OUT = !dsb_model_nout<M>::has_out && (N % 2 == 0);

#if defined(__CUDA_ARCH__)
    for (int i = (int)threadIdx.x; i < N; i += (int)blockDim.x) satol[i] = meta.atol[i];
    if (lane == 0) dsb_mbar_init(sbar);
    __syncthreads();                    // the only block-wide synchronisation: the warps are independent from here on
#else
    for (int i = 0; i < N; ++i) satol[i] = meta.atol[i];
#endif
    unsigned bar_parity = 0;

    // ---- controller (uniform over the warp: every lane holds the same values) -----------------------------------------
    int state = L_FETCH;
    int64_t inst = 0;
    int order = 1, n_equal_steps = 0;
    double t = 0.0, h = 0.0, c = 0.0, t_predict = 0.0;
    bool has_tstop = false, has_prev_error = false, jacobian_is_stale = true;
    double tstop = 0.0, prev_error_norm = 0.0;
    LaneJacobianUpdate ju; ju.init(1.0);
    LaneConvergence conv;
    conv.tol = pa.opt.nonlinear_solver_tolerance; conv.max_iter = pa.opt.max_nonlinear_solver_iterations;
    conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
    LaneStats st; st.clear();
    double pl[NP > 0 ? NP : 1];
#pragma unroll
    for (int j = 0; j < (NP > 0 ? NP : 1); ++j) pl[j] = 0.0;
    bool convergence_fail = false, newton_ok = false, first = true, reached = false, accepted = false;
    bool repredict = true, pending_etf = false, rs_ignore_small = false;
    int old_num_error_test_failures = 0, col = 0;
    double safety = 0.0, error_norm = 0.0;
    int after_rescale = L_JAC, after_jac = L_TSTOP, jac_kind = DSB_CHECKPOINT;
    double rescale_factor = 1.0;
    int fin_status = DSB_STATUS_OK;
    int stores_in_flight = 0;           // bulk-async stores whose shared-memory source may still be read (lane 0's groups)
    int stage_next = 0;                 // output staging alternates between the dead work vectors DL and YC
    auto finish = [&](int status) { fin_status = status; state = L_FINISH; };
    // the output staging vectors are about to be written by the integrator again
    auto drain_stores = [&]() {
        if (stores_in_flight > 0) {
            if (lane == 0) dsb_bulk_store_wait_read<0>();
            dsb_wsync();
            stores_in_flight = 0;
        }
    };
    // bdf.rs:694-731
    auto handle_tstop = [&](double ts) -> int {
        const double troundoff = 100.0 * eps * (dsb_abs(t) + dsb_abs(h));
        if (dsb_abs(t - ts) <= troundoff) { has_tstop = false; return 1; }
        if ((h > 0.0 && ts < t - troundoff) || (h < 0.0 && ts > t + troundoff)) {
            has_tstop = false;
            return -DSB_STATUS_STOP_TIME_BEFORE_CURRENT;
        }
        if ((h > 0.0 && t + h > ts + troundoff) || (h < 0.0 && t + h < ts - troundoff)) {
            rescale_factor = DSB_DIV(ts - t, h);
            return 2;
        }
        return 0;
    };
    // runge_kutta.rs:1313-1335
    auto pi_controller_raw = [&](double err, int eff_order) -> double {
        const double order_f = (double)eff_order;
        const double ki = DSB_DIV(pa.opt.pi_control_integral, order_f);
        const bool p_only = pa.opt.pi_control_proportional == 0.0 || !has_prev_error;
        const double kp = p_only ? 0.0 : DSB_DIV(pa.opt.pi_control_proportional, order_f);
        double v = dsb_pow(err, p_only ? -ki : -(ki + kp));
        if (!p_only) v = v * dsb_pow(prev_error_norm, kp);
        return v;
    };
    // sum of the N squared terms parked in DL, in index order (vector/nalgebra_serial.rs:395-408): every lane adds them
    // all, so the result is uniform; the loads do not depend on the sum and run ahead of it
    auto sum_terms = [&]() -> double {
        dsb_wsync();
        double acc = 0.0;
#if defined(__CUDA_ARCH__)
        const double2* const terms = reinterpret_cast<const double2*>(sm + Lay::O_DL);     // two terms per load
#pragma unroll 8
        for (int i = 0; i < N / 2; ++i) { const double2 v = terms[i]; acc += v.x; acc += v.y; }
        if (N & 1) acc += SDL(N - 1);
#else
        for (int i = 0; i < N; ++i) acc += SDL(i);
#endif
        dsb_wsync();                    // DL is free again
        return DSB_DIV(acc, (double)N);
    };
    // ||x||^2_w(ref): the terms lane-parallel into DL (x may be DL itself), then the in-order sum
    auto dcol = [&](int j) -> const double* { return Lay::D_SHARED ? sm + Lay::O_D + j * NS : gslot + Lay::G_D + j * NS; };
    // x: a shared-memory vector, or (GLOBAL_X) a column of the difference array in the warp's global-memory slot
    auto weighted_norm = [&](auto GLOBAL_X, const double* x, const double* ref) -> double {
        dsb_warp_for<DCHUNK, double>(lane, N, [&](int i) { return decltype(GLOBAL_X)::value ? dsb_ld_l2(x + i) : x[i]; }, [&](int i, double xi) {
            const double term = DSB_DIV(xi, dsb_abs(ref[i]) * pa.rtol + satol[i]);
            SDL(i) = term * term;
        });
        return sum_terms();
    };
    // the time factors of interpolate (bdf.rs:767-782, 1080-1106)
    auto time_factors = [&](double tq, double (&tf)[DSB_MAX_ORDER]) {
        double time_factor = 1.0;
#pragma unroll
        for (int j = 0; j < DSB_MAX_ORDER; ++j) {
            if (j < order) {
                const double j_t = (double)j;
                time_factor *= DSB_DIV(tq - (t - h * j_t), h * (1.0 + j_t));
            }
            tf[j] = time_factor;
        }
    };
    // columns 0 .. order of the difference array for component i (the loads of one chunk go out together)
    auto load_cols = [&](int i) -> WCols {
        WCols c;
#pragma unroll
        for (int j = 0; j <= DSB_MAX_ORDER; ++j) c.v[j] = (j <= order) ? SD_LD(j, i) : 0.0;
        return c;
    };
    auto interpolate_cols = [&](const WCols& c, const double (&tf)[DSB_MAX_ORDER]) -> double {
        double yo = c.v[0];
#pragma unroll
        for (int j = 0; j < DSB_MAX_ORDER; ++j) if (j < order) yo = tf[j] * c.v[j + 1] + yo;
        return yo;
    };
    auto interpolate_i = [&](int i, const double (&tf)[DSB_MAX_ORDER]) -> double { return interpolate_cols(load_cols(i), tf); };
    // interpolate(tq) into a shared-memory vector, one component per lane
    auto interpolate_to = [&](double tq, double* dst) {
        double tf[DSB_MAX_ORDER];
        time_factors(tq, tf);
        dsb_warp_for<DCHUNK, WCols>(lane, N, load_cols, [&](int i, const WCols& c) { dst[i] = interpolate_cols(c, tf); });
        dsb_wsync();
    };
    // the same for the output and root functions: only the components they read when the equations declare them
    auto interpolate_for_functions = [&](double tq) {
        if constexpr (NDEP > 0) {
            double tf[DSB_MAX_ORDER];
            time_factors(tq, tf);
            for (int q = lane; q < NDEP; q += LANES) { const int i = M::dep(q); SDL(i) = interpolate_i(i, tf); }
            dsb_wsync();
        } else {
            interpolate_to(tq, sm + Lay::O_DL);
        }
    };
    LaneRootFinder<(NR > 0 ? NR : 1), DsbDivShared> rf;
    rf.t0 = 0.0;
    int root_found = -1;
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;

    // one column of the solve_dense result (dense_write_out, method.rs:822-848), instance-major
    auto write_column = [&](double tq, int column) {
        double* const dst = ys_im + ((int64_t)inst * nt + column) * NOUT;
        if constexpr (dsb_model_nout<M>::has_out) {
            interpolate_for_functions(tq);
            double o[NOUT];
            M::out(vDL, pl, tq, o);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NOUT; ++k) dst[k] = o[k];
            }
            dsb_wsync();                // DL is read by every lane above and rewritten by the next column
        } else if constexpr (BULK_OUT) {
            // stage the column in a dead work vector; it leaves through a bulk-async store while the warp goes on
            if (stores_in_flight >= 2) {
                if (lane == 0) dsb_bulk_store_wait_read<1>();
                dsb_wsync();
                stores_in_flight = 1;
            }
            double* const stage = sm + (stage_next == 0 ? Lay::O_DL : Lay::O_YC);
            stage_next = stage_next == 1 ? 0 : stage_next + 1;
            interpolate_to(tq, stage);
            if (lane == 0) dsb_bulk_store(dst, stage, N);
            stores_in_flight += 1;
        } else {
            double tf[DSB_MAX_ORDER];
            time_factors(tq, tf);
            dsb_warp_for<DCHUNK, WCols>(lane, N, load_cols, [&](int i, const WCols& c) { dst[i] = interpolate_cols(c, tf); });
        }
    };

    while (true) {
        if (state == L_IDLE) break;
        // ================= FINISH =================================================================================
        if (state == L_FINISH) {
            if (lane == 0) {
                dsb_bulk_store_wait_all();
                bb.status[inst] = fin_status;
                bb.fin_t[inst] = t; bb.fin_h[inst] = h; bb.fin_order[inst] = order;
#pragma unroll
                for (int k = 0; k < DSB_NSTATS; ++k) bb.stats[(int64_t)k * B + inst] = st.v[k];
                if (NR > 0) { bb.ncols[inst] = col; bb.root_idx[inst] = root_found; }
            }
            dsb_wsync();
            stores_in_flight = 0;
            state = L_FETCH;
        }
        // ================= FETCH: next instance; new_without_initialise, set_step_size, Bdf::_new part 1 ==============
        if (state == L_FETCH) {
            unsigned long long got = 0;
            if (lane == 0) got = atomicAdd(work_counter, 1ull);
            inst = (int64_t)dsb_wbcast(got, 0);
            if (inst >= B) {
                state = L_IDLE;
            } else if (!M::HAS_MASS || bb.status[inst] == DSB_STATUS_OK) {     // else: consistent initialisation failed, keep its status
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = bb.params[(int64_t)j * B + inst];
                st.clear();
                t = pa.t0;
                if constexpr (M::HAS_MASS) {
                    // singular mass: y, dy after set_consistent and the counters so far come from dsb_band_init_kernel
                    // (state.rs:84-162)
#pragma unroll
                    for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = bb.stats[(int64_t)k * B + inst];
                    WFOR(i) { SY(i) = bb.y0[(int64_t)i * B + inst]; SD_ST(1, i, bb.dy0[(int64_t)i * B + inst]); }
                    dsb_wsync();
                } else {
                    // y = init(p, t0); dy = f(y, t0)     (state.rs:1086-1124); dy is kept in D[1] until h is known
                    WFOR(i) SY(i) = M::init_i(i, pl, pa.t0);
                    dsb_wsync();
                    WFOR(i) SD_ST(1, i, M::rhs_i(i, vY, pl, pa.t0));
                    dsb_wsync();
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                }
                // set_step_size (state.rs:1209-1277), solver order 1
                {
                    const bool is_neg_h = pa.h0 < 0.0;
                    const double d0 = dsb_sqrt(weighted_norm(std::false_type{}, sm + Lay::O_Y, sm + Lay::O_Y));
                    const double d1 = dsb_sqrt(weighted_norm(std::bool_constant<!Lay::D_SHARED>{}, dcol(1), sm + Lay::O_Y));
                    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * DSB_DIV(d0, d1);
                    WFOR(i) { const double dy = SD_LD(1, i); SYC(i) = is_neg_h ? (dy * (-h0) + SY(i)) : (dy * h0 + SY(i)); }
                    dsb_wsync();
                    const double t1 = is_neg_h ? pa.t0 - h0 : pa.t0 + h0;
                    WFOR(i) SDL(i) = M::rhs_i(i, vYC, pl, t1) - SD_LD(1, i);
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                    const double d2 = DSB_DIV(dsb_sqrt(weighted_norm(std::false_type{}, sm + Lay::O_DL, sm + Lay::O_Y)), dsb_abs(h0));
                    double max_d = d2;
                    if (max_d < d1) max_d = d1;
                    double h1;
                    if (max_d < 1e-15) { h1 = h0 * 1e-3; if (h1 < 1e-6) h1 = 1e-6; }
                    else h1 = dsb_pow(DSB_DIV(0.01, max_d), DSB_DIV(1.0, 1.0 + 1.0));
                    h = 100.0 * h0;
                    if (h > h1) h = h1;
                    if (is_neg_h) h = -h;
                }
                // state.set_problem (bdf_state.rs:72-78): D[:, 0] = y, D[:, 1] = h dy, the rest zero
                WFOR(i) {
                    SD_ST(0, i, SY(i)); SD_ST(1, i, SD_LD(1, i) * h);
#pragma unroll
                    for (int j = 2; j < DSB_NDIFF; ++j) SD_ST(j, i, 0.0);
                }
                dsb_wsync();
                order = 1; n_equal_steps = 0;
                conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
                c = h * pa.tab.alpha[1];
                jacobian_is_stale = true;
                ju.init(1.0);                                   // jacobian_update.rs:27 -- h_at_last starts at ONE
                has_tstop = false; tstop = 0.0; has_prev_error = false; prev_error_norm = 0.0;
                convergence_fail = false; first = true; reached = false; pending_etf = false; col = 0;
                t_predict = t;
                root_found = -1;
                if constexpr (NR > 0) {                         // Bdf::_new: root_finder.init(root_fn, state.y, state.t)
                    M::root(vY, pl, t, rf.g0);
                    rf.t0 = t;
                }
                jac_kind = DSB_KIND_CONSTRUCT;
                state = L_JAC;
            }
        }
        // ================= SELECT (bdf.rs:1489-1563, 1431-1442) ===========================================================
        if (state == L_SELECT) {
            const int ord = order;
            const double inf = dsb_from_bits(0x7ff0000000000000ULL);
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll 1
            for (int q = 0; q < 3; ++q) {
                if (accepted || q == 1) {
                    double err = error_norm;
                    if (q != 1) {
                        err = inf;
                        if ((q == 0) ? (ord > 1) : (ord < DSB_MAX_ORDER)) {
                            const double e = weighted_norm(std::bool_constant<!Lay::D_SHARED>{}, dcol(ord + q), sm + Lay::O_Y) * pa.tab.error_const2[ord - 1 + q];
                            err = (0.0 < e) ? e : 0.0;
                        }
                    }
                    const double v = pi_controller_raw(err, ord + q);
                    if (q == 0) f0 = v; else if (q == 1) f1 = v; else f2 = v;
                }
            }
            if (accepted) {
                int max_index = 0;                      // Iterator::max_by keeps the LAST maximum
                double fmax = f0;
                if (!(fmax > f1)) { max_index = 1; fmax = f1; }
                if (!(fmax > f2)) { max_index = 2; fmax = f2; }
                order = ord + (max_index - 1);
                double factor = safety * fmax;
                if (factor > pa.opt.max_timestep_growth) factor = pa.opt.max_timestep_growth;
                if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
                state = L_TSTOP;
                if (factor >= pa.opt.min_timestep_growth || factor <= pa.opt.max_timestep_shrink || max_index != 1) {
                    rescale_factor = factor; rs_ignore_small = false;
                    state = L_RESCALE; after_rescale = L_JAC;
                    jac_kind = DSB_STEP_SUCCESS; after_jac = L_TSTOP;
                }
            } else {
                double factor = safety * f1;
                has_prev_error = false;
                if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
                rescale_factor = factor; rs_ignore_small = false;
                state = L_RESCALE; after_rescale = L_JAC;
                jac_kind = DSB_ERROR_TEST_FAIL; after_jac = L_PREDICT;
                repredict = true; pending_etf = true;
            }
        }

        // ================= RESCALE: _update_step_size(factor) (bdf.rs:508-577) ==============================================
        // R U (rows / columns 1..k; row and column 0 are those of the identity) is built row by row by one lane and parked
        // in shared memory, then D[:, 1..k] <- D[:, 1..k] (R U) one component per lane.
        if (state == L_RESCALE) {
            const double factor = rescale_factor;
            const double new_h = factor * h;
            n_equal_steps = 0;
            const int k = order;
            const double* __restrict__ u = pa.tab.u[DSB_MAX_ORDER];         // leading dimension 6
            if (lane == 0) {
                double rrow[DSB_MAX_ORDER + 1];
#pragma unroll
                for (int l = 1; l <= DSB_MAX_ORDER; ++l) rrow[l] = 1.0;
#pragma unroll 1
                for (int i = 1; i <= k; ++i) {
                    const double i_t = (double)i;
#pragma unroll
                    for (int l = 1; l <= DSB_MAX_ORDER; ++l) rrow[l] = DSB_DIV(rrow[l] * (i_t - 1.0 - factor * (double)l), i_t);
#pragma unroll
                    for (int j = 1; j <= DSB_MAX_ORDER; ++j) {
                        double ru_ij = rrow[1] * u[j * 6 + 1];
#pragma unroll
                        for (int l = 2; l <= j; ++l) ru_ij = rrow[l] * u[j * 6 + l] + ru_ij;
                        SRU(i, j) = ru_ij;
                    }
                }
            }
            dsb_wsync();
            dsb_warp_for<DCHUNK, WCols>(lane, N, load_cols, [&](int s, const WCols& dc) {
                double nd[DSB_MAX_ORDER + 1];
#pragma unroll
                for (int j = 1; j <= DSB_MAX_ORDER; ++j) nd[j] = -0.0;      // (-0.0) + x == x: the first term is assigned
#pragma unroll
                for (int i = 1; i <= DSB_MAX_ORDER; ++i) {
                    if (i <= k) {
                        const double di = dc.v[i];
#pragma unroll
                        for (int j = 1; j <= DSB_MAX_ORDER; ++j) nd[j] = di * SRU(i, j) + nd[j];
                    }
                }
#pragma unroll
                for (int j = 1; j <= DSB_MAX_ORDER; ++j) if (j <= k) SD_ST(j, s, nd[j]);
            });
            dsb_wsync();
            c = new_h * pa.tab.alpha[k];
            h = new_h;
            conv.eta = pa.tab.eta_reset_timestep;
            if (!rs_ignore_small && dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
            else state = after_rescale;
        }

        // ================= JAC: _jacobian_updates(c, kind) / Bdf::_new's reset_jacobian =====================================
        if (state == L_JAC) {
            bool do_factor = false;
            if (jac_kind == DSB_KIND_CONSTRUCT) {
                do_factor = true;
                st.v[DSB_STAT_LINEAR_SOLVER_SETUPS] += 1;
                st.v[DSB_STAT_SETUPS_FROM_CHECKPOINT] += 1;
                after_jac = L_TSTOP;
            } else if (ju.check_rhs_jacobian_update<DsbDivShared>(pa.opt, c, jac_kind)) {
                jacobian_is_stale = true;
                ju.update_rhs_jacobian(c);
                ju.update_jacobian(c);
                do_factor = true;
            } else if (ju.check_jacobian_update<DsbDivShared>(pa.opt, c, jac_kind)) {
                ju.update_jacobian(c);
                do_factor = true;
            }
            if (do_factor) {
                if (jac_kind != DSB_KIND_CONSTRUCT) {
                    conv.eta = pa.tab.eta_reset;
                    st.record_linear_solver_setup(jac_kind);
                }
                if (jacobian_is_stale) {
                    // df/dy at (state.y, state.t) (quirk Q6), one jac_mul per colour (jacobian/mod.rs:236-256; without
                    // colouring the host supplies one colour per column: op/nonlinear_op.rs:211-220), scattered through
                    // the sparsity pattern into the band rows of the warp's global-memory slot
                    st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
                    for (int e = lane; e < LDJ * NS; e += LANES) dsb_st_l2(&gslot[Lay::G_J + e], 0.0);
                    dsb_wsync();
                    const bool one_colour_per_column = pa.ncolors == N;
#pragma unroll 1
                    for (int cc = 0; cc < pa.ncolors; ++cc) {
                        const BandColourSeed seed{meta.colmeta, cc};
                        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
                        // rows outside the band of the colour's only column hold exact zeros and are not evaluated
                        const int i0 = one_colour_per_column ? (cc - KU < 0 ? 0 : cc - KU) : 0;
                        const int i1 = one_colour_per_column ? (cc + KL > N - 1 ? N - 1 : cc + KL) : N - 1;
                        for (int i = i0 + lane; i <= i1; i += LANES) {
                            const double val = M::jac_mul_i(i, vY, pl, t, seed);
#pragma unroll
                            for (int d = -KL; d <= KU; ++d) {               // column j = i + d
                                const int j = i + d;
                                if (j >= 0 && j < N) {
                                    const int32_t m = meta.colmeta[j];
                                    if ((m & 0xffff) == cc && ((m >> (16 + KU - d)) & 1)) dsb_st_l2(&GJ(j, KU - d), val);
                                }
                            }
                        }
                    }
                    if constexpr (M::HAS_MASS) {
                        // mass.matrix_inplace(t) with the Jacobian (op/bdf.rs:273-300): column j = M e_j, beta = 0
                        WFOR(j) {
                            const BandUnitVec ej{j};
#pragma unroll
                            for (int r = 0; r < LDJ; ++r) {
                                const int i = j + r - KU;
                                dsb_st_l2(&GM(j, r), (i >= 0 && i < N) ? M::mass_i(i, ej, pl, t, 0.0, 0.0) : 0.0);
                            }
                        }
                    }
                    dsb_wsync();
                    jacobian_is_stale = false;
                }
                // df/dy comes back from the warp's slot by one bulk-async copy, straight into the band rows KL .. LDAB-1
                // of the factor storage (where A's entries of the same diagonals go)
                if (lane == 0) dsb_bulk_load(&SAB(0, KL), &GJ(0, 0), LDJ * NS, sbar);
                // A = M - c J (op/bdf.rs:282-298: J * (-c) + M) in band storage with kl extra rows for the fill-in
                const double mc = -c;
                double mrow[M::HAS_MASS ? LDJ : 1][(N + LANES - 1) / LANES];
                if constexpr (M::HAS_MASS) {             // the mass rows travel through registers while the copy is in flight
#pragma unroll
                    for (int r = 0; r < LDJ; ++r)
#pragma unroll
                        for (int q = 0; q < (N + LANES - 1) / LANES; ++q) {
                            const int j = lane + q * LANES;
                            mrow[r][q] = (j < N) ? dsb_ld_l2(&GM(j, r)) : 0.0;
                        }
                }
                dsb_mbar_wait(sbar, bar_parity);
                bar_parity ^= 1u;
#pragma unroll
                for (int q = 0; q < (N + LANES - 1) / LANES; ++q) {
                    const int j = lane + q * LANES;
                    if (j < N) {
#pragma unroll
                        for (int r = 0; r < LDAB; ++r) {
                            const int i = j + r - KV;
                            double v = 0.0;
                            if (r >= KL && i >= 0 && i < N) {
                                if constexpr (M::HAS_MASS) v = SAB(j, r) * mc + mrow[r - KL][q];
                                else v = SAB(j, r) * mc + ((i == j) ? 1.0 : 0.0);
                            }
                            SAB(j, r) = v;
                        }
                    }
                }
                dsb_wsync();
                // band LU, dgbtf2 convention, one lane
                if (lane == 0) sflag[0] = BLU::factor(sm + Lay::O_AB, spiv, sm + Lay::O_RCP, sflag + 1, sflag + 2);
                dsb_wsync();
            }
            state = after_jac;
        }

        // ================= TSTOP ==========================================================================================
        if (state == L_TSTOP) {
            bool stopped_on_root = false;
            if constexpr (NR > 0) {
                // check for a root within the accepted step (bdf.rs:1566-1579), after the step-size update and before the
                // stop time is handled; the interpolated state of the secant iteration goes to the (free) Newton residual
                if (!first) {
                    double t_root = t;
                    stopped_on_root = rf.check_root(t, [&](double (&gv)[NR]) { M::root(vY, pl, t, gv); },
                                                    [&](double t_mid, double (&gv)[NR]) {
                                                        dsb_wsync();
                                                        interpolate_for_functions(t_mid);
                                                        M::root(vDL, pl, t_mid, gv);
                                                    }, t_root, root_found);
                    dsb_wsync();
                    if (stopped_on_root) {
                        if (!free_running) {
                            // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root)
                            // (bdf.rs:1228-1262), then the state at the root in the next column (method.rs:493-503) and the
                            // end of the solve
                            while (col < nt && bb.t_eval[col] <= t_root) {
                                write_column(bb.t_eval[col], col);
                                ++col;
                            }
                            if (col < nt) {
                                write_column(t_root, col);
                                ++col;
                            }
                            t = t_root;
                        } else {
                            // the step() / interpolate() loop of the reference's harness (ode_solver/mod.rs:132-141) returns
                            // interpolate(t_root) for the point it was stepping towards and ends
                            if (col < nt) {
                                write_column(t_root, col);
                                ++col;
                            }
                        }
                        finish(DSB_STATUS_OK);
                    }
                }
            }
            int next = first ? L_PREDICT : L_OUTPUT;
            int r = 0;
            bool check = has_tstop && !stopped_on_root;
            if (first) {
                check = !free_running;
                if (free_running) next = L_OUTPUT;
                else { has_tstop = true; tstop = bb.t_eval[nt - 1]; }
            }
            if (check) {
                r = handle_tstop(tstop);
                if (r == 1) {
                    if (first) r = -DSB_STATUS_STOP_TIME_AT_CURRENT;
                    else reached = true;
                }
            }
            if (stopped_on_root) {
                // the warp is on its way to FINISH
            } else if (r < 0) {
                finish(-r);
            } else if (r == 2) {
                rs_ignore_small = true;            // "step size too small" is ignored here (bdf.rs:726-728)
                state = L_RESCALE; after_rescale = next;
            } else {
                state = next;
            }
            if (first && state != L_FETCH) {       // start of the first step()
                old_num_error_test_failures = st.v[DSB_STAT_ERROR_TEST_FAILURES];
                convergence_fail = false; repredict = true;
            }
            first = false;
        }

        // ================= OUTPUT: dense output at every t_eval passed (method.rs:761-764, 822-848) =========================
        if (state == L_OUTPUT) {
            int status = DSB_STATUS_OK;
            if constexpr (dsb_model_nout<M>::has_out && NDEP > 0 && NOUT == 1) {
                // output function with declared dependencies: up to LANES pending columns at once, one per lane -- each lane
                // interpolates the NDEP components at ITS time and evaluates the function on them; consecutive columns
                // of an instance are consecutive words of the instance-major result
                while (!free_running && h > 0.0 && col < nt) {
                    const int mycol = col + lane;
                    const double tq = mycol < nt ? bb.t_eval[mycol] : 0.0;
                    const bool mine = mycol < nt && tq <= t;
                    const unsigned ready = dsb_wballot(mine);
                    // t_eval is increasing: the ready columns are a prefix
                    int nready = 0;
                    while (nready < LANES && ((ready >> nready) & 1u)) ++nready;
                    if (nready == 0) break;
                    if (lane < nready) {
                        double tf[DSB_MAX_ORDER];
                        time_factors(tq, tf);
                        WDepVec<M, NDEP> yq;
#pragma unroll
                        for (int q = 0; q < NDEP; ++q) yq.v[q] = interpolate_i(M::dep(q), tf);
                        double o[1];
                        M::out(yq, pl, tq, o);
                        ys_im[(int64_t)inst * nt + mycol] = o[0];
                    }
                    col += nready;
                    if (nready < LANES) break;
                }
                dsb_wsync();
            }
            while (col < nt) {
                const double tq = bb.t_eval[col];
                if (free_running ? (dsb_abs(t) < dsb_abs(tq)) : !(tq <= t)) break;
                const bool is_forward = h > 0.0;
                if ((is_forward && tq > t) || (!is_forward && tq < t)) { status = DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT; break; }
                write_column(tq, col);
                ++col;
            }
            if (status != DSB_STATUS_OK) finish(status);
            else if (free_running ? (col >= nt) : reached) finish(DSB_STATUS_OK);
            else {                                  // start of the next step()
                old_num_error_test_failures = st.v[DSB_STAT_ERROR_TEST_FAILURES];
                convergence_fail = false; repredict = true;
                state = L_PREDICT;
            }
        }

        // ================= PREDICT: _predict_forward + start of a Newton solve ===============================================
        if (state == L_PREDICT) {
            drain_stores();
            if (repredict) {
                const int ord = order;
                const double a = pa.tab.alpha[ord];
                dsb_warp_for<DCHUNK, WCols>(lane, N, load_cols, [&](int i, const WCols& dc) {
                    double yp = 0.0;
                    double ps = 0.0;
#pragma unroll
                    for (int j = 0; j <= DSB_MAX_ORDER; ++j) {
                        if (j <= ord) {
                            const double d = dc.v[j];
                            yp += d;
                            if (j == 1) ps = pa.tab.gamma[1] * d;
                            else if (j >= 2) ps = pa.tab.gamma[j] * d + ps;
                        }
                    }
                    ps *= a;
                    ps -= yp;
                    SYP_ST(i, yp); SPSI_ST(i, ps); SYC(i) = yp;
                });
                t_predict = t + h;
            } else {
                dsb_warp_for<VCHUNK, double>(lane, N, [&](int i) { return SYP_LD(i); }, [&](int i, double v) { SYC(i) = v; });
            }
            dsb_wsync();
            state = L_NEWTON;
            if (pending_etf) {
                pending_etf = false;
                st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
                if (st.v[DSB_STAT_ERROR_TEST_FAILURES] - old_num_error_test_failures >= pa.opt.max_error_test_failures)
                    finish(DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES);
            }
            conv.reset();
        }

        // ================= NEWTON: one iteration (newton.rs:13-36, line_search.rs:48-69) =====================================
        if (state == L_NEWTON) {
            // delta = F(y) = M (y + psi - y0) - c f(t, y)   (op/bdf.rs:240-256)
            const double mc = -c;
            auto residual = [&]() {
                dsb_warp_for<VCHUNK, double>(lane, N, [&](int i) { return M::HAS_MASS ? 0.0 : SPSI_LD(i); }, [&](int i, double psi) {
                    const double f = M::rhs_i(i, vYC, pl, t_predict);
                    if constexpr (M::HAS_MASS) SDL(i) = M::mass_i(i, vTMP, pl, t_predict, mc, f);    // gemv_inplace(x, t, beta, y): y = M x + beta y
                    else SDL(i) = (SYC(i) + psi) + mc * f;
                });
                dsb_wsync();
            };
            residual();
            st.v[DSB_STAT_RHS_CALLS] += 1;
            const bool ok = sflag[0] == 0;                      // else LaError::LuSolveFailed: the factors hold a zero pivot
            if (ok) {
                int rc = 1;
                const int nsw = sflag[1];
                if (lane == 0) {
                    double* const ab = sm + Lay::O_AB; double* const rc_ = sm + Lay::O_RCP; double* const rhs = sm + Lay::O_DL;
                    rc = nsw == 0 ? BLU::template solve<false, 0>(ab, spiv, rc_, rhs, sflag + 2, 0)
                       : nsw <= BLU::MAXSW ? BLU::template solve<false, 1>(ab, spiv, rc_, rhs, sflag + 2, nsw)
                                           : BLU::template solve<false, 2>(ab, spiv, rc_, rhs, sflag + 2, nsw);
                }
                dsb_wsync();
                rc = dsb_wbcast(rc, 0);
                if (rc == 2 || pa.reserved1 != 0) {
                    // a quotient of the fast back substitution could not be vouched for (or the test hook asks for this path):
                    // the same right-hand side again, through the plain IEEE divisions
                    residual();
                    if (lane == 0) BLU::template solve<true, 2>(sm + Lay::O_AB, spiv, sm + Lay::O_RCP, sm + Lay::O_DL, sflag + 2, nsw);
                    dsb_wsync();
                }
            }
            if (!ok) {
                newton_ok = false; state = L_POST;              // LuSolveFailed
            } else {
                dsb_warp_for<VCHUNK, double>(lane, N, [&](int i) { return SYP_LD(i); }, [&](int i, double yp) {
                    const double dl = SDL(i);
                    SYC(i) = SYC(i) - dl;
                    // Newton norm weights use the PREDICTOR (line_search.rs:67, convergence.rs:64-66)
                    const double term = DSB_DIV(dl, dsb_abs(yp) * pa.rtol + satol[i]);
                    SDL(i) = term * term;
                });
                const double norm = dsb_sqrt(sum_terms());
                // Convergence::check_new_iteration (convergence.rs:68-139)
                conv.niter += 1;
                const bool have_rate = conv.has_old_norm;
                double px, py;
                if (have_rate) { px = DSB_DIV(norm, conv.old_norm); py = DSB_DIV(1.0, (double)(conv.niter - 1)); }
                else { const double min_eta = 1e4 * eps; px = (conv.eta < min_eta) ? min_eta : conv.eta; py = 0.8; }
                const double pw = dsb_pow(px, py);
                int s = LANE_CONTINUE;
                if (have_rate) {
                    const double rate = pw;
                    if (rate > 0.9) s = LANE_DIVERGED;
                    else if (DSB_DIV(dsb_powi(rate, conv.max_iter - conv.niter), 1.0 - rate) * norm > conv.tol) s = LANE_DIVERGED;
                    else conv.eta = DSB_DIV(rate, 1.0 - rate);
                } else {
                    conv.eta = pw;
                }
                if (s != LANE_DIVERGED && conv.eta * norm < conv.tol) s = LANE_CONVERGED;
                if (conv.niter == 1) { conv.has_old_norm = true; conv.old_norm = norm; }   // frozen at the FIRST norm (quirk Q3)
                if (s == LANE_CONVERGED) { newton_ok = true; state = L_POST; }
                else if (s == LANE_DIVERGED || conv.niter >= conv.max_iter) { newton_ok = false; state = L_POST; }
            }
        }
        // ================= POST: a Newton solve ended (bdf.rs:1338-1563) =====================================================
        if (state == L_POST) {
            st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
            if (newton_ok) {
                const int ord = order;
                {   // error_control: ||d||^2_w(state.y) * error_const2[order - 1], d = y - y_predict
                    dsb_warp_for<VCHUNK, double>(lane, N, [&](int i) { return SYP_LD(i); }, [&](int i, double yp) {
                        const double d = SYC(i) - yp;
                        const double term = DSB_DIV(d, dsb_abs(SY(i)) * pa.rtol + satol[i]);
                        SDL(i) = term * term;
                    });
                    const double err = sum_terms() * pa.tab.error_const2[ord - 1];
                    error_norm = (0.0 < err) ? err : 0.0;
                }
                const double maxiter = (double)conv.max_iter;
                const double niter = (double)conv.niter;
                safety = DSB_DIV(0.9 * (2.0 * maxiter + 1.0), 2.0 * maxiter + niter);
                if (error_norm <= 1.0) {
                    // ---- accepted: _update_diff, state.y <- PREDICTOR (quirk Q1) ----
                    dsb_warp_for<DCHUNK, WColsYp>(lane, N, [&](int i) {
                        WColsYp r;
                        r.c = load_cols(i);
                        r.c.v[DSB_MAX_ORDER + 1] = SD_LD(ord + 1, i);       // the old D[:, ord + 1]
                        r.yp = SYP_LD(i);
                        return r;
                    }, [&](int i, const WColsYp& r) {
                        const WCols& dc = r.c;
                        const double yp = r.yp;
                        const double d = SYC(i) - yp;
                        double above = d;                                   // the new D[:, ord + 1]
                        SD_ST(ord + 2, i, d - dc.v[DSB_MAX_ORDER + 1]);
                        SD_ST(ord + 1, i, d);
#pragma unroll
                        for (int j = DSB_MAX_ORDER; j >= 0; --j) {
                            if (j <= ord) {
                                above = dc.v[j] + 1.0 * above;
                                SD_ST(j, i, above);
                            }
                        }
                        SY(i) = yp;
                    });
                    dsb_wsync();
                    t = t_predict;
                    st.v[DSB_STAT_STEPS] += 1;
                    ju.step();
                    has_prev_error = true; prev_error_norm = error_norm;
                    n_equal_steps += 1;
                    accepted = true;
                    state = (n_equal_steps > ord) ? L_SELECT : L_TSTOP;
                } else {
                    accepted = false;
                    state = L_SELECT;
                }
            } else {
                // ---- Newton failed ----
                st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] += 1;
                has_prev_error = false;
                if (st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] > pa.opt.max_nonlinear_solver_failures) {
                    finish(DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES);
                } else if (convergence_fail) {
                    rescale_factor = 0.3; rs_ignore_small = false;
                    state = L_RESCALE; after_rescale = L_JAC;
                    jac_kind = DSB_SECOND_CONVERGENCE_FAIL; after_jac = L_PREDICT;
                    repredict = true;
                } else {
                    convergence_fail = true;
                    state = L_JAC; jac_kind = DSB_FIRST_CONVERGENCE_FAIL; after_jac = L_PREDICT;
                    repredict = false;                          // retry from the SAME predictor
                }
            }
        }
    }
#undef SD_LD
#undef SD_ST
#undef SY
#undef SYP_LD
#undef SYP_ST
#undef SYC
#undef SPSI_LD
#undef SPSI_ST
#undef SDL
#undef SAB
#undef SRU
#undef GJ
#undef GM
#undef DSB_DIV
#undef WFOR
}
