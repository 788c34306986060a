// dsb_band_lu.cuh -- per-lane band LU of the banded one-thread-per-instance kernels (dsb_band_bdf_kernel.cuh,
// dsb_band_sdirk_kernel.cuh, dsb_band_init_kernel.cuh).
//
// The lane's matrices live in its global-memory column (word w of the lane at g[w * LS], see
// dsb_band_bdf_kernel.cuh); the iteration matrix is held in LAPACK band storage (dgbtf2 convention, 2 kl + ku + 1
// rows per column: entry (i, j) at word o_ab + j * LDAB + kl + ku + i - j, room for the fill-in of partial pivoting)
// and factored sequentially by the lane.  The substitutions keep the running right-hand-side entries in a REGISTER
// WINDOW (kl + 1 resp. kl + ku + 1 values), so the recurrences never wait for a global-memory round trip.
//
// Same arithmetic as nalgebra 0.35 `DMatrix::lu()` / `LU::solve_mut` as called from
// diffsol-la/src/linear_solver/nalgebra/lu.rs:31-51 (first maximum as pivot, reciprocal-pivot scaling,
// `a = (-u) * l + a` in ascending pivot order, column-axpy substitutions): every operation that is skipped has an
// exactly zero multiplier or pivot-row entry, the interchanges are interleaved with the forward substitution as in
// dsb_coop.cuh:warp_band_solve, and the solutions are bit-identical to the dense path.
#pragma once
#include "dsb_lane.cuh"

// Blocked component loop: store(i, compute(i)) for i in [0, n), with the loads and arithmetic of U consecutive
// components issued BEFORE any of their stores.  Every vector of a lane lives behind the same base pointer with a
// run-time stride, so the compiler cannot prove that the store of component i does not alias the loads of component
// i + 1 and would otherwise serialise the loop at one global-memory round trip per component (a warp issues in
// order: it stalls at the store until the loads that feed it return).  compute() must not read what store() writes
// for another component.
template <int U, class R, class C, class S>
DSB_DEV void band_for(const int n, C&& compute, S&& store) {
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += U) {
        R r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u < n) r[u] = compute(i0 + u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u < n) store(i0 + u, r[u]);
    }
}

template <int N, int KL, int KU, class DIV, int U = 2>      // U: unroll factor of the substitution loops
struct LaneBandLU {
    static constexpr int KV = KL + KU, LDAB = 2 * KL + KU + 1;
    static_assert(KL >= 1 && KL <= 2 && KU >= 1 && KU <= 2, "register windows are sized for kl, ku <= 2");

    // P A = L U in place; pivot offsets (row j interchanged with row j + piv[j]) as doubles at o_piv
    static DSB_DEV void factor(double* __restrict__ g, const size_t LS, const int o_ab, const int o_piv) {
#define GAB_(j, r) g[(size_t)(o_ab + (j) * LDAB + (r)) * LS]
#define GPIV_(j) g[(size_t)(o_piv + (j)) * LS]
        int jlast = 0;                                   // last column touched by the fill-in so far
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const int km = (KL < N - 1 - j) ? KL : (N - 1 - j);
            double colv[KL + 1];
#pragma unroll
            for (int d = 0; d <= KL; ++d) colv[d] = (d <= km) ? GAB_(j, KV + d) : 0.0;
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
            if (diag == 0.0) { GPIV_(j) = 0.0; continue; }
            GPIV_(j) = (double)jp;
            { const int cand = (j + KU + jp < N - 1) ? (j + KU + jp) : (N - 1); if (cand > jlast) jlast = cand; }
            if (jp != 0) {
#pragma unroll
                for (int q = 0; q <= KV; ++q) {              // columns j .. jlast (at most kv + 1 of them)
                    const int cq = j + q;
                    if (cq <= jlast) {
                        const double a = GAB_(cq, KV - q), b = GAB_(cq, KV - q + jp);
                        GAB_(cq, KV - q) = b; GAB_(cq, KV - q + jp) = a;
                    }
                }
#pragma unroll
                for (int d = 0; d <= KL; ++d) {              // the register copy of column j follows the interchange
                    const double a = colv[0];
                    if (jp == d && d != 0) { colv[0] = colv[d]; colv[d] = a; }
                }
            }
            if (km > 0) {
                const double inv_diag = 1.0 / colv[0];
#pragma unroll
                for (int d = 1; d <= KL; ++d) if (d <= km) { colv[d] *= inv_diag; GAB_(j, KV + d) = colv[d]; }
#pragma unroll
                for (int q = 1; q <= KV; ++q) {              // columns j + 1 .. jlast
                    const int cq = j + q;
                    if (cq <= jlast) {
                        const double mpk = -GAB_(cq, KV - q);
#pragma unroll
                        for (int d = 1; d <= KL; ++d)
                            if (d <= km) GAB_(cq, KV - q + d) = mpk * colv[d] + GAB_(cq, KV - q + d);
                    }
                }
            }
        }
    }

    // b <- A^-1 b for the vector at o_b; false when a zero pivot is met (LaError::LuSolveFailed): b is then partly solved.
    // Both sweeps run in blocks of U rows: the block's pivots, multipliers and incoming right-hand-side entries are
    // loaded first, the recurrence runs in registers, the block's results are stored last (see band_for).
    static DSB_DEV bool solve(double* __restrict__ g, const size_t LS, const int o_ab, const int o_piv, const int o_b) {
#define GB_(i) g[(size_t)(o_b + (i)) * LS]
        // forward substitution with the interchanges interleaved; b[j .. j + kl] travels in registers
        {
            double w[KL + 1];
#pragma unroll
            for (int d = 0; d <= KL; ++d) w[d] = GB_(d);
#pragma unroll 1
            for (int j0 = 0; j0 + 1 < N; j0 += U) {
                int jp[U];
                double lm_[U][KL], bin[U], out[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u;
                    if (j + 1 < N) {
                        jp[u] = (int)GPIV_(j);
#pragma unroll
                        for (int d = 1; d <= KL; ++d) lm_[u][d - 1] = (j + d < N) ? GAB_(j, KV + d) : 0.0;
                        bin[u] = (j + 1 + KL < N) ? GB_(j + 1 + KL) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u;
                    if (j + 1 < N) {
                        if (jp[u] != 0) {
                            const double a = w[0];
#pragma unroll
                            for (int d = 1; d <= KL; ++d) if (jp[u] == d) { w[0] = w[d]; w[d] = a; }
                        }
                        const double bj = w[0];
                        out[u] = bj;
                        const double nbj = -bj;
#pragma unroll
                        for (int d = 1; d <= KL; ++d) if (j + d < N) w[d] = nbj * lm_[u][d - 1] + w[d];
#pragma unroll
                        for (int d = 0; d < KL; ++d) w[d] = w[d + 1];
                        w[KL] = bin[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) if (j0 + u + 1 < N) GB_(j0 + u) = out[u];
            }
            GB_(N - 1) = w[0];
        }
        // back substitution, column-axpy form; b[i - kv .. i] travels in registers
        bool ok = true;
        {
            double w[KV + 1];
#pragma unroll
            for (int e = 0; e <= KV; ++e) w[e] = GB_(N - 1 - e);
#pragma unroll 1
            for (int i0 = N - 1; i0 >= 0; i0 -= U) {
                double up[U][KV + 1], bin[U], out[U];
                bool have[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 - u;
                    if (i >= 0) {
#pragma unroll
                        for (int e = 0; e <= KV; ++e) up[u][e] = (i - e >= 0) ? GAB_(i, KV - e) : 0.0;
                        bin[u] = (i - 1 - KV >= 0) ? GB_(i - 1 - KV) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 - u;
                    have[u] = false;
                    if (i >= 0) {
                        const double diag = up[u][0];
                        if (diag == 0.0) ok = false;
                        if (ok) {
                            const double coeff = DIV::div(w[0], diag);
                            out[u] = coeff; have[u] = true;
                            const double ncoeff = -coeff;
#pragma unroll
                            for (int e = 1; e <= KV; ++e) if (i - e >= 0) w[e] = ncoeff * up[u][e] + w[e];
                        }
#pragma unroll
                        for (int e = 0; e < KV; ++e) w[e] = w[e + 1];
                        w[KV] = bin[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) if (have[u]) GB_(i0 - u) = out[u];
            }
        }
        return ok;
#undef GB_
#undef GAB_
#undef GPIV_
    }
};
