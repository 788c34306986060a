// dsb_coop_bdf_kernel.cuh -- the path for systems too large for one thread per instance (n > 16) that are not banded:
// ONE THREAD BLOCK per instance, one thread per state component.  One kernel, two instantiations over the same
// cooperative pieces: Bdf::step (RK = false; the file's name is from when it was the only one) and Sdirk::step
// (RK = true: TR-BDF2 / ESDIRK34 from the tableau in pa.rk).  Control flow is uniform over the block
// (every thread carries the controller scalars and takes the same decisions from the same shared-memory
// data), so the reference's nested loops are kept as they are; what is parallel is every vector / matrix
// operation inside them.  Vectors (the difference array D, y, predictor, Newton iterate, ...) live in shared
// memory; J, M and the LU factors (n^2 doubles each: 512 KB at n = 256) live in global memory and are
// streamed through the blocked LU of dsb_coop.cuh.  Blocks are persistent and draw instances from a global
// work counter.
//
// Bit-exactness: element-wise operations are order-free; everything the reference sums sequentially
// (squared_norm: nalgebra_serial.rs:395-408; LU; substitutions) is summed in the same order here (the norm's
// terms are computed in parallel and added up by one thread).
//
// Restated functions: the same lists as dsb_bdf_kernel.cuh, dsb_sdirk_kernel.cuh and dsb_init_kernel.cuh (Bdf::_new,
// Bdf::step / Rk::_new, Sdirk::step and everything they call, new_and_consistent with InitOp, set_step_size,
// solve_dense with its RootFound / reset branches, dense_write_out with an output function), paths relative to
// /root/reference/crates/diffsol/src.
#pragma once
#include <type_traits>

#include "dsb_coop.cuh"
#include "dsb_roots.cuh"
#include "dsb_lane.cuh"
#include "dsb_models.h"


// Component-wise view of the equations.  Models written component-wise (dsb_models.h: `*_i`) are used
// directly; for the small whole-vector models the component is picked out of a full evaluation (O(n) per
// component: only used to run the reference's small test problems through this kernel).
template <class M, bool CW = dsb_is_componentwise<M>::value> struct CoopEval;
template <class M> struct CoopEval<M, true> {
    static __device__ __forceinline__ double rhs_i(int i, const double* x, const double* p, double t) { return M::rhs_i(i, x, p, t); }
    template <class V>
    static __device__ __forceinline__ double jac_mul_i(int i, const double* x, const double* p, double t, const V& v) { return M::jac_mul_i(i, x, p, t, v); }
    static __device__ __forceinline__ double mass_i(int i, const double* x, const double* p, double t, double beta, double yi) { return M::mass_i(i, x, p, t, beta, yi); }
    static __device__ __forceinline__ double init_i(int i, const double* p, double t) { return M::init_i(i, p, t); }
};
template <class M> struct CoopEval<M, false> {
    static __device__ double rhs_i(int i, const double* x, const double* p, double t) {
        double xl[M::N], yl[M::N];
        for (int k = 0; k < M::N; ++k) { xl[k] = x[k]; yl[k] = 0.0; }
        M::rhs(xl, p, t, yl);
        return yl[i];
    }
    template <class V>
    static __device__ double jac_mul_i(int i, const double* x, const double* p, double t, const V& v) {
        double xl[M::N], vl[M::N], yl[M::N];
        for (int k = 0; k < M::N; ++k) { xl[k] = x[k]; vl[k] = v[k]; yl[k] = 0.0; }
        M::jac_mul(xl, p, t, vl, yl);
        return yl[i];
    }
    static __device__ double mass_i(int i, const double* x, const double* p, double t, double beta, double yi) {
        double xl[M::N], yl[M::N];
        for (int k = 0; k < M::N; ++k) { xl[k] = x[k]; yl[k] = 0.0; }
        yl[i] = yi;
        M::mass(xl, p, t, beta, yl);
        return yl[i];
    }
    static __device__ double init_i(int i, const double* p, double t) {
        double yl[M::N];
        M::init(p, t, yl);
        return yl[i];
    }
};

// unit vector e_j as an indexable object (seed of one Jacobian / mass column)
struct CoopUnitVec {
    int j;
    __device__ __forceinline__ double operator[](int k) const { return k == j ? 1.0 : 0.0; }
};

// Global-memory workspace of one resident block (instance in flight)
struct DsbCoopWorkspace {
    double* jac;      // [nblocks][n*n]  df/dy
    double* mass;     // [nblocks][n*n]  M (DAE only)
    double* lu;       // [nblocks][n*n]  factors of M - cJ
    int32_t* piv;     // [nblocks][n]
    const double* atol;   // [n]
    const int32_t* color; // [n] colour of every column (-1: empty column), or nullptr: dense assembly
    const uint8_t* nz;    // [n*n] column-major sparsity pattern of df/dy
};

template <class M>
struct CoopBdfLayout {
    static constexpr int N = M::N;
    static constexpr int NVEC = DSB_NDIFF + 9;          // D[8], y, yp, ycur, psi, dlt, tmp, scr, atol, dy
    static constexpr int THREADS = (N + 31) / 32 * 32 > 128 ? 128 : (N + 31) / 32 * 32;
    static size_t smem_bytes() { return (size_t)NVEC * N * sizeof(double) + coop_lu_smem_bytes_host(N) + 64 + (size_t)N * sizeof(int); }
};

// Resident blocks the register allocation is asked to allow: DSB_COOP_TARGET_THREADS lanes per SM (the kernel is
// latency bound -- barriers, global-memory factors -- so resident warps matter more than registers per lane).
// Measured on SPM (n = 42, 64-lane blocks): 4 -> 16 resident blocks per SM = 1.75x, spills included; at n = 256 shared
// memory allows 2 blocks either way.
#ifndef DSB_COOP_TARGET_THREADS
#define DSB_COOP_TARGET_THREADS(threads) ((threads) <= 64 ? 1024 : 256)
#endif
// RK = false: Bdf; RK = true: Sdirk (TR-BDF2 / ESDIRK34, the tableau in pa.rk) on the same cooperative pieces
template <class M, bool RK = false>
__global__ void __launch_bounds__(CoopBdfLayout<M>::THREADS, DSB_COOP_TARGET_THREADS(CoopBdfLayout<M>::THREADS) / CoopBdfLayout<M>::THREADS)
dsb_coop_bdf_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa, const __grid_constant__ DsbBatchBuffers bb,
                                const __grid_constant__ DsbCoopWorkspace ws, unsigned long long* __restrict__ work_counter) {
    constexpr int N = M::N;
    constexpr int NP = M::NP;
    constexpr int NR = dsb_model_nroots<M>::value;
    constexpr int NOUT = dsb_model_nout<M>::value;
    constexpr bool HAS_OUT = dsb_model_nout<M>::has_out;
    typedef CoopEval<M> E;
    extern __shared__ unsigned char dsb_coop_bdf_smem[];
    double* const vec = (double*)dsb_coop_bdf_smem;
    double* const Dm = vec;                              // D[j][i] at Dm[j * N + i]
    double* const ys = vec + DSB_NDIFF * N;              // state.y
    double* const yp = ys + N;                           // y_predict
    double* const yc = yp + N;                           // Newton iterate / y_delta
    double* const psi = yc + N;                          // psi - y_predict
    double* const dlt = psi + N;                         // Newton residual / update
    double* const tmpv = dlt + N;
    double* const scr = tmpv + N;                        // per-component terms of a norm
    double* const atolv = scr + N;
    double* const dys = atolv + N;                       // state.dy (initialisation only)
    const CoopScratch sc = coop_carve(dys + N, N);
    int* const piv_s = sc.bcast + 4;                     // pivots of the shared-memory band factors (tridiagonal fast path)
    __shared__ int s_band[2];                            // kl, ku of the current J / M
    __shared__ int s_banded;                             // 1: the current factors are the shared-memory band factors
    __shared__ double s_red;                             // broadcast of a sequential reduction
    __shared__ long long s_inst;
    __shared__ double s_p[NP > 0 ? NP : 1];

    const int tid = threadIdx.x, T = blockDim.x;
    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const double eps = 2.220446049250313e-16;
    double* const Jg = ws.jac + (size_t)blockIdx.x * N * N;
    double* const Mg = M::HAS_MASS ? ws.mass + (size_t)blockIdx.x * N * N : nullptr;
    double* const LUg = ws.lu + (size_t)blockIdx.x * N * N;
    int32_t* const pivg = ws.piv + (size_t)blockIdx.x * N;

    for (int i = tid; i < N; i += T) atolv[i] = ws.atol[i];
    if (ws.color != nullptr) {                             // coloured assembly only ever writes pattern entries
        for (int e = tid; e < N * N; e += T) Jg[e] = 0.0;
    }

    // sum_i term_i / n with term_i = (x_i / (|y_i| rtol + atol_i))^2, terms in parallel, the sum sequential
    auto squared_norm = [&](const double* x, const double* yref) -> double {
        __syncthreads();
        for (int i = tid; i < N; i += T) {
            const double term = x[i] / (dsb_abs(yref[i]) * pa.rtol + atolv[i]);
            scr[i] = term * term;
        }
        __syncthreads();
        if (tid == 0) {
            double acc = 0.0;
            for (int i = 0; i < N; ++i) acc += scr[i];
            s_red = acc / (double)N;
        }
        __syncthreads();
        return s_red;
    };

    // dense path: LU of LUg (global memory, blocked); banded path: see reset_jacobian
    auto lu_factor = [&]() {
        __syncthreads();
        if (tid == 0) s_banded = 0;
        coop_lu_factor(LUg, N, pivg, sc);
        __syncthreads();
    };
    auto lu_solve = [&](double* b) -> bool {
        if (s_banded) {
            __syncthreads();
            if (s_band[0] == 1 && s_band[1] == 1) { if (tid == 0) sc.bcast[2] = thread_band_solve<1, 1>(sc.panel, N, piv_s, b) ? 1 : 0; }
            else if (tid < 32) { const bool ok = warp_band_solve(sc.panel, N, s_band[0], s_band[1], pivg, b); if (tid == 0) sc.bcast[2] = ok ? 1 : 0; }
            __syncthreads();
            return sc.bcast[2] != 0;
        }
        return coop_lu_solve(LUg, N, pivg, b, sc);
    };

    while (true) {
        __syncthreads();
        if (tid == 0) s_inst = (long long)atomicAdd(work_counter, 1ull);
        __syncthreads();
        const int64_t inst = s_inst;
        if (inst >= B) break;

        LaneStats st; st.clear();
        if (tid < NP) s_p[tid] = bb.params[(int64_t)tid * B + inst];
        __syncthreads();
        const double* p = s_p;
        int status = DSB_STATUS_OK;

        // ================= new_and_consistent (state.rs:969-997, 1086-1124) =================
        double t = pa.t0, h = pa.h0;
        for (int i = tid; i < N; i += T) ys[i] = E::init_i(i, p, pa.t0);
        __syncthreads();
        for (int i = tid; i < N; i += T) dys[i] = E::rhs_i(i, ys, p, pa.t0);
        st.v[DSB_STAT_RHS_CALLS] += 1;
        __syncthreads();

        // df/dy at (x, tt) into Jg: one jac_mul per column (op/nonlinear_op.rs:211-220) or, with colouring, one per
        // colour (jacobian/mod.rs:236-256: seed every column of the colour, scatter the product through the pattern)
        auto eval_jacobian = [&](const double* x, double tt) {
            st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
            if (ws.color != nullptr) {
                st.v[DSB_STAT_RHS_JAC_MULS] += pa.ncolors;
                for (int c = 0; c < pa.ncolors; ++c) {
                    __syncthreads();
                    for (int k = tid; k < N; k += T) tmpv[k] = (ws.color[k] == c) ? 1.0 : 0.0;
                    __syncthreads();
                    for (int i = tid; i < N; i += T) {
                        const double ci = E::jac_mul_i(i, x, p, tt, tmpv);
                        for (int j = 0; j < N; ++j)
                            if (ws.color[j] == c && ws.nz[(size_t)j * N + i]) Jg[(size_t)j * N + i] = ci;
                    }
                }
                __syncthreads();
            } else {
                st.v[DSB_STAT_RHS_JAC_MULS] += N;
                for (int j = 0; j < N; ++j) {
                    const CoopUnitVec v{j};
                    for (int i = tid; i < N; i += T) Jg[(size_t)j * N + i] = E::jac_mul_i(i, x, p, tt, v);
                }
            }
        };
        // ---- set_consistent (state.rs:84-162, op/init.rs:14-131) ----
        if (M::HAS_MASS) {
            // mass matrix at t0: column j = M e_j, with beta = 0
            for (int j = 0; j < N; ++j) {
                __syncthreads();
                for (int i = tid; i < N; i += T) tmpv[i] = (i == j) ? 1.0 : 0.0;
                __syncthreads();
                for (int i = tid; i < N; i += T) Mg[(size_t)j * N + i] = E::mass_i(i, tmpv, p, pa.t0, 0.0, 0.0);
            }
            __syncthreads();
            // algebraic indices: zero diagonal; flags kept in scr as 0/1 is not possible (scr is norm scratch),
            // so they are recomputed from Mg where needed
            __shared__ int s_nalg;
            if (tid == 0) {
                int c = 0;
                for (int i = 0; i < N; ++i) c += (Mg[(size_t)i * N + i] == 0.0) ? 1 : 0;
                s_nalg = c;
            }
            __syncthreads();
            if (s_nalg > 0) {
                auto is_alg = [&](int i) -> bool { return Mg[(size_t)i * N + i] == 0.0; };
                eval_jacobian(ys, pa.t0);
                __syncthreads();
                // InitOp jac = (-M_u | f_v ; 0 | g_v) into LUg; neg_mass = (-M_u | 0 ; 0 | 0) overwrites Jg
                for (int e = tid; e < N * N; e += T) {
                    const int j = e / N, i = e % N;
                    double jv = 0.0, nm = 0.0;
                    if (!is_alg(j)) { if (!is_alg(i)) { const double m_u = Mg[e] * -1.0; jv = m_u; nm = m_u; } }
                    else jv = Jg[e];
                    LUg[e] = jv; Jg[e] = nm;
                }
                __syncthreads();
                // the InitOp Jacobian is constant: factor once, keep a copy is unnecessary because LU setups
                // beyond the first only re-factor the same matrix (bitwise the same factors)
                lu_factor();
                // y_tmp (yc) = (dy at differential idx, y at algebraic idx); yerr (yp) = y_tmp; y0 copy in psi
                for (int i = tid; i < N; i += T) { yc[i] = is_alg(i) ? ys[i] : dys[i]; yp[i] = yc[i]; psi[i] = ys[i]; }
                __syncthreads();
                // fun(x) -> dlt: y0[alg] = x[alg]; out = f(y0); out = neg_mass * x + out (column sweep)
                auto fun = [&](const double* x) {
                    __syncthreads();
                    for (int i = tid; i < N; i += T) if (is_alg(i)) psi[i] = x[i];
                    __syncthreads();
                    for (int i = tid; i < N; i += T) {
                        double o = E::rhs_i(i, psi, p, pa.t0);
                        for (int j = 0; j < N; ++j) o = Jg[(size_t)j * N + i] * x[j] + o;
                        dlt[i] = o;
                    }
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                    __syncthreads();
                };
                LaneConvergence conv;
                conv.tol = pa.opt.nonlinear_solver_tolerance; conv.eta = pa.tab.eta_reset;
                conv.max_iter = pa.opt.ic_max_newton_iterations; conv.old_norm = 0.0; conv.reset();
                const double tau = pa.opt.ic_step_reduction_factor, c_armijo = pa.opt.ic_armijo_constant;
                const double steptol = pa.tab.ic_steptol;
                bool ok = false;
                for (int setup = 0; setup < pa.opt.ic_max_linear_solver_setups && status == DSB_STATUS_OK && !ok; ++setup) {
                    conv.reset();
                    double ls_norm = 1.0;
                    int result = -1;
                    for (int i = tid; i < N; i += T) dlt[i] = 0.0;
                    for (int it = 0; it < conv.max_iter && result < 0; ++it) {
                        int res = LANE_CONTINUE;
                        bool have_res = false;
                        if (pa.opt.ic_use_linesearch) {
                            if (conv.niter == 0) {
                                fun(yc);
                                if (!lu_solve(dlt)) { result = 2; break; }
                                ls_norm = dsb_sqrt(squared_norm(dlt, yp));
                                if (conv.check_norm(ls_norm) == LANE_CONVERGED) {
                                    for (int i = tid; i < N; i += T) yc[i] -= dlt[i];
                                    __syncthreads();
                                    res = LANE_CONVERGED; have_res = true;
                                }
                            }
                            if (!have_res) {
                                // x0 -> tmpv, delta0 -> Dm row 0 (D is not in use yet)
                                __syncthreads();
                                for (int i = tid; i < N; i += T) { tmpv[i] = yc[i]; Dm[i] = dlt[i]; }
                                __syncthreads();
                                const double norm = ls_norm;
                                const double phi0 = norm * norm * 0.5, two_phi0 = norm * norm;
                                const double min_alpha = steptol / norm;
                                double alpha = 1.0;
                                int ls_status = 1;
                                for (int li = 0; li < pa.opt.ic_max_linesearch_iterations; ++li) {
                                    for (int q = tid; q < N; q += T) yc[q] = (-alpha) * Dm[q] + yc[q];
                                    fun(yc);
                                    if (!lu_solve(dlt)) { ls_status = 2; break; }
                                    const double new_norm = dsb_sqrt(squared_norm(dlt, yp));
                                    const double phi1 = new_norm * new_norm * 0.5;
                                    if (phi1 <= phi0 - c_armijo * alpha * two_phi0) {
                                        ls_norm = new_norm;
                                        res = conv.check_norm(new_norm); have_res = true; ls_status = 0;
                                        break;
                                    }
                                    if (alpha < min_alpha) { ls_status = 2; break; }
                                    alpha *= tau;
                                    __syncthreads();
                                    for (int q = tid; q < N; q += T) yc[q] = tmpv[q];
                                    __syncthreads();
                                }
                                if (ls_status != 0) { result = 2; break; }
                            }
                        } else {
                            fun(yc);
                            if (!lu_solve(dlt)) { result = 2; break; }
                            for (int i = tid; i < N; i += T) yc[i] -= dlt[i];
                            res = conv.check_new_iteration(dsb_sqrt(squared_norm(dlt, yp)));
                        }
                        if (res == LANE_CONVERGED) result = 0;
                        else if (res == LANE_DIVERGED) result = 2;
                    }
                    if (result < 0) result = 1;
                    if (result == 0) ok = true;
                    else if (result == 2) status = DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
                    else {
                        __syncthreads();
                        for (int i = tid; i < N; i += T) yp[i] = yc[i];
                        __syncthreads();
                    }
                }
                if (!ok && status == DSB_STATUS_OK) status = DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
                __syncthreads();
                if (status == DSB_STATUS_OK) {
                    for (int i = tid; i < N; i += T) {
                        if (is_alg(i)) { ys[i] = yc[i]; dys[i] = 0.0; }
                        else dys[i] = yc[i];
                    }
                }
                __syncthreads();
            }
        }

        // ---- set_step_size (state.rs:1209-1277) ----
        if (status == DSB_STATUS_OK) {
            const bool is_neg_h = pa.h0 < 0.0;
            const double d0 = dsb_sqrt(squared_norm(ys, ys));
            const double d1 = dsb_sqrt(squared_norm(dys, ys));
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            __syncthreads();
            for (int i = tid; i < N; i += T) tmpv[i] = is_neg_h ? (dys[i] * (-h0) + ys[i]) : (dys[i] * h0 + ys[i]);
            __syncthreads();
            const double t1 = is_neg_h ? pa.t0 - h0 : pa.t0 + h0;
            for (int i = tid; i < N; i += T) dlt[i] = E::rhs_i(i, tmpv, p, t1) - dys[i];
            st.v[DSB_STAT_RHS_CALLS] += 1;
            const double d2 = dsb_sqrt(squared_norm(dlt, ys)) / dsb_abs(h0);
            double max_d = d2;
            if (max_d < d1) max_d = d1;
            double h1;
            if (max_d < 1e-15) { h1 = h0 * 1e-3; if (h1 < 1e-6) h1 = 1e-6; }
            else h1 = dsb_pow(0.01 / max_d, 1.0 / (1.0 + (RK ? (double)pa.rk.order : 1.0)));   // solver_order: 1 for Bdf (problem.rs:597-602), the tableau's for Sdirk
            h = 100.0 * h0;
            if (h > h1) h = h1;
            if (is_neg_h) h = -h;
        }

        // results of the instance (thread 0 writes)
        auto write_results = [&](int status_, double t_, double h_, int order_, int col_, int root_found_) {
            __syncthreads();
            if (tid == 0) {
                bb.status[inst] = status_;
                bb.fin_t[inst] = t_; bb.fin_h[inst] = h_; bb.fin_order[inst] = order_;
                for (int k = 0; k < DSB_NSTATS; ++k) bb.stats[(int64_t)k * B + inst] = st.v[k];
                if (status_ == DSB_STATUS_OK) bb.ncols[inst] = col_;
                if (NR > 0) bb.root_idx[inst] = root_found_;
            }
        };
        // A = J * mc + M (M - cJ for Bdf, M - c h J for Sdirk) and its LU: in LAPACK band storage in shared memory when the
        // band of J / M is narrow, else dense in global memory
        auto assemble_and_factor = [&](double mc) {
            const int kl = s_band[0], ku = s_band[1];
            if (pa.coop_dense_only == 0 && N > 32 && 2 * kl + ku + 1 <= 32) {
                const int kv = kl + ku;
                double* ab = sc.panel;
                __syncthreads();
                for (int e = tid; e < N * 32; e += T) {
                    const int j = e >> 5, d = e & 31;
                    const int r = j - kv + d;
                    double v = 0.0;
                    if (d <= kv + kl && r >= 0 && r < N && r >= j - ku) {
                        const double m_ji = M::HAS_MASS ? Mg[(size_t)j * N + r] : ((r == j) ? 1.0 : 0.0);
                        v = Jg[(size_t)j * N + r] * mc + m_ji;
                    }
                    ab[e] = v;
                }
                __syncthreads();
                if (kl == 1 && ku == 1) { if (tid == 0) thread_band_factor<1, 1>(ab, N, piv_s); }
                else if (tid < 32) warp_band_factor(ab, N, kl, ku, pivg);
                if (tid == 0) s_banded = 1;
                __syncthreads();
            } else {
                for (int e = tid; e < N * N; e += T) {
                    const int j = e / N, i = e % N;
                    const double m_ji = M::HAS_MASS ? Mg[e] : ((i == j) ? 1.0 : 0.0);
                    LUg[e] = Jg[e] * mc + m_ji;
                }
                __syncthreads();
                lu_factor();
            }
        };
        // M at time tt into Mg: column j = M e_j (beta = 0)
        auto eval_mass = [&](double tt) {
            for (int j = 0; j < N; ++j) {
                __syncthreads();
                for (int i = tid; i < N; i += T) tmpv[i] = (i == j) ? 1.0 : 0.0;
                __syncthreads();
                for (int i = tid; i < N; i += T) Mg[(size_t)j * N + i] = E::mass_i(i, tmpv, p, tt, 0.0, 0.0);
            }
        };

        // apply_reset (state.rs:246-270, no mass matrix): y <- reset(x, t) with x the state at the root (a shared vector),
        // dy <- f(y, t).  Component-wise equations are evaluated cooperatively; the small whole-vector ones by every
        // thread on local copies.
        auto apply_reset_at = [&](const double* x, double tt) {
            if constexpr (dsb_model_has_reset<M>::value) {
                if constexpr (dsb_is_componentwise<M>::value) {
                    __syncthreads();
                    for (int i = tid; i < N; i += T) ys[i] = M::reset_i(i, x, p, tt);
                    __syncthreads();
                    for (int i = tid; i < N; i += T) dys[i] = E::rhs_i(i, ys, p, tt);
                    __syncthreads();
                } else {
                    double yl[N], yr[N], dyr[N];
                    for (int i = 0; i < N; ++i) yl[i] = x[i];
                    M::reset(yl, p, tt, yr);
                    M::rhs(yr, p, tt, dyr);
                    __syncthreads();
                    for (int i = tid; i < N; i += T) { ys[i] = yr[i]; dys[i] = dyr[i]; }
                    __syncthreads();
                }
                st.v[DSB_STAT_RHS_CALLS] += 1;
            }
        };

        if constexpr (RK) {
        // ================= Rk::_new + Sdirk::_new (runge_kutta.rs:100-190, sdirk.rs:80-160) =================
        const int ns = pa.rk.s;
        const double cg = pa.rk.a[1 * ns + 1];                       // the SDIRK diagonal
        const int start = (pa.rk.a[0] == 0.0) ? 1 : 0;               // explicit first stage (ESDIRK, TR-BDF2)
        double* const diff = Dm;                                     // diff[j][i] = h k_j at diff[j * N + i], j < 4
        double* const oy = Dm + 4 * N;                               // old_state.y (last stage value / previous step)
        double* const phi = Dm + 5 * N;                              // SdirkCallable.phi
        double* const errv = Dm + 6 * N;                             // embedded error estimate
        double* const xc = yc;                                       // stage iterate (old_state.dy)
        double old_t = t, op_h = h;
        bool has_tstop = false, has_prev_error = false, jacobian_is_stale = true, is_jacobian_set = false;
        double tstop = 0.0, prev_error_norm = 0.0;
        LaneJacobianUpdate ju; ju.init(1.0); ju.update_jacobian(h); ju.update_rhs_jacobian(h);
        LaneConvergence conv;
        conv.tol = pa.opt.nonlinear_solver_tolerance; conv.max_iter = pa.opt.max_nonlinear_solver_iterations;
        conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
        LaneRootFinder<(NR > 0 ? NR : 1), DsbDivInline> rf;
        rf.t0 = t;
        for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
        int root_found = -1;
        int col = 0;

        // SdirkCallable::jacobian_inplace (op/sdirk.rs:257-292) + set_linearisation: J at phi + c x with x = state.y (phi is
        // whatever the last stage left there), M, both when stale; A = M - (c h) J; LU
        auto reset_jacobian = [&](double tt) {
            __syncthreads();
            if (jacobian_is_stale) {
                for (int i = tid; i < N; i += T) psi[i] = cg * ys[i] + phi[i];
                __syncthreads();
                eval_jacobian(psi, tt);
                if (M::HAS_MASS) eval_mass(tt);
                jacobian_is_stale = false;
                __syncthreads();
                coop_band_scan(Jg, M::HAS_MASS ? Mg : nullptr, N, s_band);
            }
            assemble_and_factor(-(cg * op_h));
            is_jacobian_set = true;
        };
        // Sdirk::jacobian_updates (sdirk.rs:256-304)
        auto jacobian_updates = [&](double hh, int kind) {
            bool did_update = false;
            if (ju.check_rhs_jacobian_update(pa.opt, hh, kind)) {
                jacobian_is_stale = true;
                reset_jacobian(t);
                ju.update_rhs_jacobian(hh);
                ju.update_jacobian(hh);
                conv.eta = pa.tab.eta_reset;
                did_update = true;
            } else if (ju.check_jacobian_update(pa.opt, hh, kind)) {
                reset_jacobian(t);
                ju.update_jacobian(hh);
                conv.eta = pa.tab.eta_reset;
                did_update = true;
            }
            if (did_update) st.record_linear_solver_setup(kind);
        };
        // runge_kutta.rs:752-781.  0 = nothing, 1 = TstopReached, < 0 = -status
        auto handle_tstop = [&](double ts) -> int {
            const double troundoff = 100.0 * eps * (dsb_abs(t) + dsb_abs(h));
            if (dsb_abs(t - ts) <= troundoff) return 1;
            if ((h > 0.0 && ts < t - troundoff) || (h < 0.0 && ts > t + troundoff)) return -DSB_STATUS_STOP_TIME_BEFORE_CURRENT;
            if ((h > 0.0 && t + h > ts + troundoff) || (h < 0.0 && t + h < ts - troundoff)) {
                const double f = (ts - t) / h;
                h *= f;
            }
            return 0;
        };
        // interpolate_inplace (runge_kutta.rs:1080-1127; :962-981 beta dense output, :1004-1024 Hermite) on [old_t, t]
        auto interpolate_i = [&](double theta, int i) -> double {
            if (pa.rk.has_beta) {
                const double th2 = theta * theta;
                double yo = oy[i];
                for (int j = 0; j < ns; ++j) {
                    double bf = pa.rk.beta[j] * theta;
                    bf = pa.rk.beta[ns + j] * th2 + bf;
                    yo = diff[j * N + i] * bf + yo;
                }
                return yo;
            }
            const double al1 = theta - 1.0, be1 = 1.0 - 2.0 * theta;
            const double al2 = 1.0 - theta, be2 = theta * (theta - 1.0);
            const double u0 = oy[i], u1 = ys[i];
            double v = u1;
            v -= u0;
            v = al1 * diff[i] + be1 * v;
            v = theta * diff[(ns - 1) * N + i] + v;
            v = al2 * u0 + be2 * v;
            v = theta * u1 + v;
            return v;
        };
        auto interpolate_to_shared = [&](double tq, double* dst) {
            const double dt = t - old_t;
            const double theta = (dt == 0.0) ? 1.0 : (tq - old_t) / dt;
            __syncthreads();
            for (int i = tid; i < N; i += T) dst[i] = interpolate_i(theta, i);
            __syncthreads();
        };
        auto interpolate_and_write = [&](double tq, int column) -> int {
            const bool is_forward = h > 0.0;
            if ((is_forward && (tq > t || tq < old_t)) || (!is_forward && (tq < t || tq > old_t)))
                return DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT;
            if constexpr (HAS_OUT) {
                interpolate_to_shared(tq, tmpv);
                if (tid == 0) {
                    double o[NOUT];
                    M::out(tmpv, p, tq, o);
                    for (int k = 0; k < NOUT; ++k) bb.ys[((int64_t)column * NOUT + k) * B + inst] = o[k];
                }
                return DSB_STATUS_OK;
            }
            const double dt = t - old_t;
            const double theta = (dt == 0.0) ? 1.0 : (tq - old_t) / dt;
            __syncthreads();
            for (int i = tid; i < N; i += T) bb.ys[((int64_t)column * N + i) * B + inst] = interpolate_i(theta, i);
            return DSB_STATUS_OK;
        };

        if (status == DSB_STATUS_OK) {
            if (ws.color != nullptr) {                     // consistent initialisation used Jg as scratch
                __syncthreads();
                for (int e = tid; e < N * N; e += T) Jg[e] = 0.0;
                __syncthreads();
            }
            for (int i = tid; i < N; i += T) {
                for (int j = 0; j < DSB_NDIFF; ++j) Dm[j * N + i] = 0.0;
                oy[i] = ys[i];
            }
            __syncthreads();
            if constexpr (NR > 0) { M::root(ys, p, t, rf.g0); rf.t0 = t; }      // Rk::_new: root_finder.init(root_fn, y, t)
            if (!free_running) {
                has_tstop = true; tstop = bb.t_eval[nt - 1];
                const int r = handle_tstop(tstop);
                if (r == 1) { has_tstop = false; status = DSB_STATUS_STOP_TIME_AT_CURRENT; }
                else if (r < 0) status = -r;
            }
        }

        // ================= solve_dense loop (method.rs:721-818) around Sdirk::step (sdirk.rs:409-543) =================
        while (status == DSB_STATUS_OK && col < nt) {
            if (free_running) {
                while (col < nt && !(dsb_abs(t) < dsb_abs(bb.t_eval[col]))) {
                    const int e = interpolate_and_write(bb.t_eval[col], col);
                    if (e) { status = e; break; }
                    ++col;
                }
                if (col >= nt || status != DSB_STATUS_OK) break;
            }
            int step_result = 0;                                    // 0 internal, 1 tstop reached, 2 root found
            {
                double hs = h;                                      // rk.start_step()
                if (dsb_abs(hs) < pa.opt.min_timestep) { status = DSB_STATUS_STEP_SIZE_TOO_SMALL; break; }
                op_h = hs;
                int nattempts = 0;
                bool updated_jacobian = false;
                double factor = 1.0, error_norm = 0.0;
                while (true) {
                    if (start == 1) {                               // start_step_attempt (runge_kutta.rs:505-535)
                        __syncthreads();
                        for (int k = tid; k < N; k += T) diff[k] = hs * dys[k];
                        __syncthreads();
                    }
                    bool failed = false;
                    for (int i = start; i < ns; ++i) {
                        // Rk::do_stage_sdirk (runge_kutta.rs:631-750): set_phi, predict_stage_sdirk, Newton on
                        // F(x) = M x - h f(phi + c x)
                        const double t_stage = t + pa.rk.c[i] * hs;
                        double al = 0.0, be = 0.0;
                        if (i >= 2) {
                            const double cc = (pa.rk.c[i] - pa.rk.c[i - 2]) / (pa.rk.c[i - 1] - pa.rk.c[i - 2]);
                            al = -cc; be = 1.0 + cc;
                        }
                        __syncthreads();
                        for (int k = tid; k < N; k += T) {
                            double ph = ys[k];
                            for (int j = 0; j < i; ++j) ph = diff[j * N + k] * pa.rk.a[j * ns + i] + ph;
                            phi[k] = ph;
                            if (i == 0) xc[k] = hs * dys[k];
                            else if (i == 1) xc[k] = diff[k];
                            else xc[k] = al * diff[(i - 2) * N + k] + be * diff[(i - 1) * N + k];
                        }
                        __syncthreads();
                        if (!is_jacobian_set) {                     // the lazy first reset_jacobian (runge_kutta.rs:661-665)
                            reset_jacobian(t_stage);
                            st.record_linear_solver_setup(DSB_CHECKPOINT);
                        }
                        bool ok = false;
                        conv.reset();
                        for (int it = 0; it < conv.max_iter; ++it) {
                            __syncthreads();
                            for (int k = tid; k < N; k += T) tmpv[k] = cg * xc[k] + phi[k];
                            __syncthreads();
                            for (int k = tid; k < N; k += T) dlt[k] = E::rhs_i(k, tmpv, p, t_stage);
                            st.v[DSB_STAT_RHS_CALLS] += 1;
                            __syncthreads();
                            const double beta = -op_h;
                            for (int k = tid; k < N; k += T)
                                dlt[k] = M::HAS_MASS ? E::mass_i(k, xc, p, t_stage, beta, dlt[k]) : (xc[k] + beta * dlt[k]);
                            __syncthreads();
                            if (!lu_solve(dlt)) break;
                            for (int k = tid; k < N; k += T) xc[k] -= dlt[k];
                            const double norm = dsb_sqrt(squared_norm(dlt, ys));
                            const int sres = conv.check_new_iteration(norm);
                            if (sres == LANE_CONVERGED) { ok = true; break; }
                            if (sres == LANE_DIVERGED) break;
                        }
                        st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
                        if (!ok) { failed = true; break; }
                        __syncthreads();
                        for (int k = tid; k < N; k += T) {
                            oy[k] = cg * xc[k] + phi[k];            // get_f_eval: the stage value
                            diff[i * N + k] = xc[k];
                        }
                        __syncthreads();
                    }
                    if (failed) {                                   // sdirk.rs:436-472
                        if (!updated_jacobian) {
                            updated_jacobian = true;
                            jacobian_updates(hs, DSB_FIRST_CONVERGENCE_FAIL);
                        } else {
                            hs *= 0.3;
                            conv.eta = pa.tab.eta_reset_timestep;
                            op_h = hs;
                            jacobian_updates(hs, DSB_SECOND_CONVERGENCE_FAIL);
                        }
                        has_prev_error = false;
                        st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] += 1;
                        if (st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] > pa.opt.max_nonlinear_solver_failures) { status = DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES; break; }
                        if (dsb_abs(hs) < pa.opt.min_timestep) { status = DSB_STATUS_STEP_SIZE_TOO_SMALL; break; }
                        continue;
                    }
                    // rk.error_norm (runge_kutta.rs:783-800): error = diff . d ; [error = M error] ; error = LU^-1 error
                    __syncthreads();
                    for (int k = tid; k < N; k += T) {
                        double e = diff[k] * pa.rk.d[0];
                        for (int j = 1; j < ns; ++j) e = diff[j * N + k] * pa.rk.d[j] + e;
                        errv[k] = e;
                    }
                    __syncthreads();
                    double* ev = errv;
                    if (M::HAS_MASS) {
                        for (int k = tid; k < N; k += T) {
                            double e = Mg[k] * errv[0];
                            for (int j = 1; j < N; ++j) e = Mg[(size_t)j * N + k] * errv[j] + e;
                            dlt[k] = e;
                        }
                        __syncthreads();
                        ev = dlt;
                    }
                    if (!lu_solve(ev)) { status = DSB_STATUS_LU_SOLVE_FAILED; break; }
                    {
                        const double e = squared_norm(ev, ys);
                        error_norm = (0.0 < e) ? e : 0.0;
                    }
                    {   // Rk::factor (runge_kutta.rs:466-495) + pi_controller_raw (:1313-1335)
                        const double maxiter = (double)conv.max_iter;
                        const double niter = (double)conv.niter;
                        const double safety = 0.9 * ((2.0 * maxiter + 1.0) / (2.0 * maxiter + niter));
                        const double order_f = (double)(pa.rk.order + 1);
                        const double ki = pa.opt.pi_control_integral / order_f;
                        double raw;
                        if (pa.opt.pi_control_proportional == 0.0 || !has_prev_error) raw = dsb_pow(error_norm, -ki);
                        else {
                            const double kp = pa.opt.pi_control_proportional / order_f;
                            raw = dsb_pow(error_norm, -(ki + kp)) * dsb_pow(prev_error_norm, kp);
                        }
                        double f = safety * raw;
                        if (f > pa.opt.max_timestep_shrink && f < pa.opt.min_timestep_growth) f = 1.0;
                        if (f < pa.opt.min_timestep_shrink) f = pa.opt.min_timestep_shrink;
                        if (f > pa.opt.max_timestep_growth) f = pa.opt.max_timestep_growth;
                        factor = f;
                    }
                    if (error_norm < 1.0) break;
                    hs *= factor;
                    conv.eta = pa.tab.eta_reset_timestep;
                    op_h = hs;
                    jacobian_updates(hs, DSB_ERROR_TEST_FAIL);
                    nattempts += 1;
                    has_prev_error = false;
                    st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
                    if (nattempts >= pa.opt.max_error_test_failures) { status = DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES; break; }
                    if (dsb_abs(hs) < pa.opt.min_timestep) { status = DSB_STATUS_STEP_SIZE_TOO_SMALL; break; }
                }
                if (status != DSB_STATUS_OK) break;
                // accepted (sdirk.rs:531-542) + Rk::step_accepted (runge_kutta.rs:894-960)
                const double new_h = hs * factor;
                if (factor != 1.0) conv.eta = pa.tab.eta_reset_timestep;
                op_h = new_h;
                jacobian_updates(new_h, DSB_STEP_SUCCESS);
                ju.step();
                has_prev_error = true; prev_error_norm = error_norm;
                const double inv_h = 1.0 / hs;
                __syncthreads();
                for (int k = tid; k < N; k += T) {
                    const double y_new = oy[k];                     // old_state.y held the last stage value
                    oy[k] = ys[k];                                  // swap: old_state <- previous state
                    ys[k] = y_new;
                    dys[k] = xc[k] * inv_h;                         // old_state.dy *= 1/h, then swapped in
                }
                __syncthreads();
                old_t = t;
                t = t + hs;
                h = new_h;
                st.v[DSB_STAT_STEPS] += 1;
                if constexpr (NR > 0) {   // also in the step()/interpolate() loop of the reference's harness (free_running), which returns interpolate(t_root) and ends (ode_solver/mod.rs:134-141)
                    // check for a root within the accepted step (runge_kutta.rs:935-948), before the stop time is handled
                    double t_root = t;
                    const bool stopped_on_root = rf.check_root(t, [&](double (&gv)[NR]) { M::root(ys, p, t, gv); },
                                                               [&](double t_mid, double (&gv)[NR]) {
                                                                   interpolate_to_shared(t_mid, yp);
                                                                   M::root(yp, p, t_mid, gv);
                                                               }, t_root, root_found);
                    if (stopped_on_root) {
                        // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root),
                        // then the state at the root in the next column (method.rs:493-503)
                        while (!free_running && col < nt && bb.t_eval[col] <= t_root) {
                            (void)interpolate_and_write(bb.t_eval[col], col);
                            ++col;
                        }
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (sdirk.rs:368-374 -> state.rs:279-306), a new
                                // stop time, then Rk::start_step (runge_kutta.rs:446-464) finds the state mutated:
                                // root finder re-initialised, stop time set again; step size, Jacobian and LU stay
                                ended = false;
                                interpolate_to_shared(t_root, yp);
                                t = t_root;
                                apply_reset_at(yp, t);
                                root_found = -1;
                                if (t < bb.t_eval[nt - 1]) {
                                    step_result = 3;
                                    has_tstop = true; tstop = bb.t_eval[nt - 1];
                                    int r = handle_tstop(tstop);                              // method.rs:792
                                    if (r == 0) {
                                        M::root(ys, p, t, rf.g0); rf.t0 = t;
                                        r = handle_tstop(tstop);                              // start_step: set_stop_time(tstop)
                                    }
                                    if (r == 1) { has_tstop = false; status = DSB_STATUS_STOP_TIME_AT_CURRENT; break; }
                                    else if (r < 0) { status = -r; break; }
                                } else {
                                    step_result = 1;                                          // TstopReached
                                }
                            }
                        }
                        if (ended) {
                            if (col < nt) {
                                (void)interpolate_and_write(t_root, col);
                                ++col;
                            }
                            __syncthreads();
                            if (!free_running) t = t_root;      // state_mut_back; the harness loop leaves the state at the end of the step
                            step_result = 2;
                        }
                    }
                }
                if (step_result == 3) step_result = 0;             // a reset was applied: the stop time is set already
                else if (has_tstop && step_result == 0) {
                    const int r = handle_tstop(tstop);
                    if (r == 1) { step_result = 1; has_tstop = false; }
                    else if (r < 0) { status = -r; break; }
                }
            }
            if (step_result == 2) break;                           // RootFound ends the solve
            if (!free_running) {
                while (col < nt && bb.t_eval[col] <= t) {
                    const int e = interpolate_and_write(bb.t_eval[col], col);
                    if (e) { status = e; break; }
                    ++col;
                }
                if (step_result == 1) break;
            }
        }
        write_results(status, t, h, pa.rk.order, col, root_found);
        } else {
        // ================= Bdf::_new (bdf.rs:230-368) =================
        int order = 1, n_equal_steps = 0;
        double c = h * pa.tab.alpha[1], t_predict = t;
        bool has_tstop = false, has_prev_error = false, jacobian_is_stale = true;
        double tstop = 0.0, prev_error_norm = 0.0;
        LaneJacobianUpdate ju; ju.init(1.0);
        LaneConvergence conv;
        conv.tol = pa.opt.nonlinear_solver_tolerance; conv.max_iter = pa.opt.max_nonlinear_solver_iterations;
        conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();

        // BdfCallable::jacobian_inplace + NalgebraLU::set_linearisation: (re)evaluate J, M when stale; A = M - cJ; LU
        auto reset_jacobian = [&]() {
            __syncthreads();
            if (jacobian_is_stale) {
                eval_jacobian(ys, t);
                if (M::HAS_MASS) eval_mass(t);
                jacobian_is_stale = false;
                __syncthreads();
                coop_band_scan(Jg, M::HAS_MASS ? Mg : nullptr, N, s_band);       // structure only changes when J / M do
            }
            assemble_and_factor(-c);
        };
        auto jacobian_updates = [&](double cc, int kind) {
            bool did_update = false;
            if (ju.check_rhs_jacobian_update(pa.opt, cc, kind)) {
                jacobian_is_stale = true;
                reset_jacobian();
                ju.update_rhs_jacobian(cc);
                ju.update_jacobian(cc);
                conv.eta = pa.tab.eta_reset;
                did_update = true;
            } else if (ju.check_jacobian_update(pa.opt, cc, kind)) {
                reset_jacobian();
                ju.update_jacobian(cc);
                conv.eta = pa.tab.eta_reset;
                did_update = true;
            }
            if (did_update) st.record_linear_solver_setup(kind);
        };
        // _update_step_size (bdf.rs:508-577): RU = R(order, factor) U; D[:, 0..=order] <- D[:, 0..=order] RU
        auto update_step_size = [&](double factor, double* new_h_out) -> int {
            const double new_h = factor * h;
            n_equal_steps = 0;
            const int nr = order + 1;
            double r[36], ru[36];
            for (int q = 0; q < nr * nr; ++q) r[q] = 0.0;
            for (int j = 0; j < nr; ++j) r[j * nr] = 1.0;
            for (int j = 1; j < nr; ++j) {
                const double j_t = (double)j;
                for (int i = 1; i < nr; ++i) {
                    const double i_t = (double)i;
                    r[j * nr + i] = r[j * nr + i - 1] * (i_t - 1.0 - factor * j_t) / i_t;
                }
            }
            const double* u = pa.tab.u[order];
            for (int j = 0; j < nr; ++j)
                for (int l = 0; l < nr; ++l) {
                    const double ulj = u[j * nr + l];
                    for (int i = 0; i < nr; ++i) {
                        if (l == 0) ru[j * nr + i] = r[l * nr + i] * ulj;
                        else ru[j * nr + i] = r[l * nr + i] * ulj + ru[j * nr + i];
                    }
                }
            __syncthreads();
            for (int i = tid; i < N; i += T) {
                double row[DSB_MAX_ORDER + 1];
                for (int j = 0; j < nr; ++j) {
                    double acc = Dm[0 * N + i] * ru[j * nr + 0];
                    for (int l = 1; l < nr; ++l) acc = Dm[l * N + i] * ru[j * nr + l] + acc;
                    row[j] = acc;
                }
                for (int j = 0; j < nr; ++j) Dm[j * N + i] = row[j];
            }
            __syncthreads();
            c = new_h * pa.tab.alpha[order];
            h = new_h;
            conv.eta = pa.tab.eta_reset_timestep;
            if (new_h_out) *new_h_out = new_h;
            if (dsb_abs(h) < pa.opt.min_timestep) return DSB_STATUS_STEP_SIZE_TOO_SMALL;
            return DSB_STATUS_OK;
        };
        auto predict_forward = [&]() {
            __syncthreads();
            for (int i = tid; i < N; i += T) {
                double a = 0.0;
                for (int j = 0; j <= order; ++j) a += Dm[j * N + i];
                double ps = pa.tab.gamma[1] * Dm[1 * N + i];
                for (int j = 2; j <= order; ++j) ps = pa.tab.gamma[j] * Dm[j * N + i] + ps;
                ps *= pa.tab.alpha[order];
                ps -= a;
                yp[i] = a; psi[i] = ps;
            }
            t_predict = t + h;
            __syncthreads();
        };
        auto handle_tstop = [&](double ts) -> int {
            const double troundoff = 100.0 * eps * (dsb_abs(t) + dsb_abs(h));
            if (dsb_abs(t - ts) <= troundoff) { has_tstop = false; return 1; }
            if ((h > 0.0 && ts < t - troundoff) || (h < 0.0 && ts > t + troundoff)) { has_tstop = false; return -DSB_STATUS_STOP_TIME_BEFORE_CURRENT; }
            if ((h > 0.0 && t + h > ts + troundoff) || (h < 0.0 && t + h < ts - troundoff)) {
                const double factor = (ts - t) / h;
                (void)update_step_size(factor, nullptr);
            }
            return 0;
        };
        auto pi_controller_raw = [&](double err, int eff_order) -> double {
            const double order_f = (double)eff_order;
            const double ki = pa.opt.pi_control_integral / order_f;
            if (pa.opt.pi_control_proportional == 0.0 || !has_prev_error) return dsb_pow(err, -ki);
            const double kp = pa.opt.pi_control_proportional / order_f;
            return dsb_pow(err, -(ki + kp)) * dsb_pow(prev_error_norm, kp);
        };
        // interpolate_from_diff (bdf.rs:767-782) into a (free) shared vector, for the output / root functions
        auto interpolate_to_shared = [&](double tq, double* dst) {
            __syncthreads();
            for (int i = tid; i < N; i += T) {
                double time_factor = 1.0;
                double yo = Dm[i];
                for (int j = 0; j < order; ++j) {
                    const double j_t = (double)j;
                    time_factor *= (tq - (t - h * j_t)) / (h * (1.0 + j_t));
                    yo = time_factor * Dm[(j + 1) * N + i] + yo;
                }
                dst[i] = yo;
            }
            __syncthreads();
        };
        // one column of the solve_dense result (dense_write_out, method.rs:822-848): the interpolated state, or -- for
        // equations with an output function -- out(y(tq), tq)
        auto interpolate_and_write = [&](double tq, int col) -> int {
            const bool is_forward = h > 0.0;
            if ((is_forward && tq > t) || (!is_forward && tq < t)) return DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT;
            if constexpr (HAS_OUT) {
                interpolate_to_shared(tq, tmpv);
                if (tid == 0) {
                    double o[NOUT];
                    M::out(tmpv, p, tq, o);
                    for (int k = 0; k < NOUT; ++k) bb.ys[((int64_t)col * NOUT + k) * B + inst] = o[k];
                }
                return DSB_STATUS_OK;
            }
            __syncthreads();
            for (int i = tid; i < N; i += T) {
                double time_factor = 1.0;
                double yo = Dm[i];
                for (int j = 0; j < order; ++j) {
                    const double j_t = (double)j;
                    time_factor *= (tq - (t - h * j_t)) / (h * (1.0 + j_t));
                    yo = time_factor * Dm[(j + 1) * N + i] + yo;
                }
                bb.ys[((int64_t)col * N + i) * B + inst] = yo;
            }
            return DSB_STATUS_OK;
        };
        // root finding (nonlinear_solver/root.rs; bdf.rs:361-366, 1566-1579): every thread of the block runs the same
        // scalar iteration on the same values (the state it reads is in shared memory), so the block's control flow
        // stays uniform; only compiled for equations with roots
        LaneRootFinder<(NR > 0 ? NR : 1), DsbDivInline> rf;
        rf.t0 = t;
        for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
        int root_found = -1;

        int col = 0;
        if (status == DSB_STATUS_OK) {
            if (ws.color != nullptr) {                     // consistent initialisation used Jg as scratch
                __syncthreads();
                for (int e = tid; e < N * N; e += T) Jg[e] = 0.0;
                __syncthreads();
            }
            reset_jacobian();
            st.v[DSB_STAT_LINEAR_SOLVER_SETUPS] += 1;
            st.v[DSB_STAT_SETUPS_FROM_CHECKPOINT] += 1;
            for (int i = tid; i < N; i += T) {
                for (int j = 0; j < DSB_NDIFF; ++j) Dm[j * N + i] = 0.0;
                Dm[i] = ys[i]; Dm[N + i] = dys[i] * h;
            }
            __syncthreads();
            if constexpr (NR > 0) { M::root(ys, p, t, rf.g0); rf.t0 = t; }      // Bdf::_new: root_finder.init(root_fn, y, t)
            if (!free_running) {
                has_tstop = true; tstop = bb.t_eval[nt - 1];
                const int r = handle_tstop(tstop);
                if (r == 1) { has_tstop = false; status = DSB_STATUS_STOP_TIME_AT_CURRENT; }
                else if (r < 0) status = -r;
            }
        }

        // ================= solve_dense loop (method.rs:721-818) around Bdf::step (bdf.rs:1277-1589) =================
        while (status == DSB_STATUS_OK && col < nt) {
            if (free_running) {
                while (col < nt && !(dsb_abs(t) < dsb_abs(bb.t_eval[col]))) {
                    const int e = interpolate_and_write(bb.t_eval[col], col);
                    if (e) { status = e; break; }
                    ++col;
                }
                if (col >= nt || status != DSB_STATUS_OK) break;
            }
            // ---- step() ----
            int step_result = 0;                                    // 0 internal, 1 tstop reached
            {
                double safety = 0.0, error_norm = 0.0, new_h = 0.0;
                const int old_etf = st.v[DSB_STAT_ERROR_TEST_FAILURES];
                bool convergence_fail = false;
                predict_forward();
                while (true) {
                    const int ord = order;
                    __syncthreads();
                    for (int i = tid; i < N; i += T) yc[i] = yp[i];
                    __syncthreads();
                    // newton_iteration + NoLineSearch
                    bool ok = false;
                    conv.reset();
                    for (int it = 0; it < conv.max_iter; ++it) {
                        __syncthreads();
                        for (int i = tid; i < N; i += T) {
                            const double f = E::rhs_i(i, yc, p, t_predict);
                            tmpv[i] = yc[i] + psi[i];
                            dlt[i] = f;
                        }
                        st.v[DSB_STAT_RHS_CALLS] += 1;
                        __syncthreads();
                        const double mc = -c;
                        for (int i = tid; i < N; i += T) {
                            dlt[i] = M::HAS_MASS ? E::mass_i(i, tmpv, p, t_predict, mc, dlt[i]) : (tmpv[i] + mc * dlt[i]);
                        }
                        __syncthreads();
                        if (!lu_solve(dlt)) break;
                        for (int i = tid; i < N; i += T) yc[i] -= dlt[i];
                        const double norm = dsb_sqrt(squared_norm(dlt, yp));
                        const int s = conv.check_new_iteration(norm);
                        if (s == LANE_CONVERGED) { ok = true; break; }
                        if (s == LANE_DIVERGED) break;
                    }
                    st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
                    if (ok) {
                        __syncthreads();
                        for (int i = tid; i < N; i += T) yc[i] -= yp[i];
                        __syncthreads();
                    } else {
                        st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] += 1;
                        if (st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] > pa.opt.max_nonlinear_solver_failures) { status = DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES; break; }
                        has_prev_error = false;
                        if (convergence_fail) {
                            const int e = update_step_size(0.3, &new_h);
                            if (e) { status = e; break; }
                            jacobian_updates(new_h * pa.tab.alpha[ord], DSB_SECOND_CONVERGENCE_FAIL);
                            predict_forward();
                        } else {
                            jacobian_updates(h * pa.tab.alpha[ord], DSB_FIRST_CONVERGENCE_FAIL);
                            convergence_fail = true;
                        }
                        continue;
                    }
                    {
                        const double err = squared_norm(yc, ys) * pa.tab.error_const2[order - 1];
                        error_norm = (0.0 < err) ? err : 0.0;
                    }
                    const double maxiter = (double)conv.max_iter;
                    const double niter = (double)conv.niter;
                    safety = 0.9 * (2.0 * maxiter + 1.0) / (2.0 * maxiter + niter);
                    if (error_norm <= 1.0) break;
                    double factor = safety * pi_controller_raw(error_norm, ord + 1);
                    has_prev_error = false;
                    if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
                    const int e = update_step_size(factor, &new_h);
                    if (e) { status = e; break; }
                    jacobian_updates(new_h * pa.tab.alpha[ord], DSB_ERROR_TEST_FAIL);
                    predict_forward();
                    st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
                    if (st.v[DSB_STAT_ERROR_TEST_FAILURES] - old_etf >= pa.opt.max_error_test_failures) { status = DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES; break; }
                }
                if (status != DSB_STATUS_OK) break;
                // accepted: _update_diff, state.y <- predictor
                __syncthreads();
                for (int i = tid; i < N; i += T) {
                    const double d = yc[i];
                    Dm[(order + 2) * N + i] = d - Dm[(order + 1) * N + i];
                    Dm[(order + 1) * N + i] = d;
                    for (int j = order; j >= 0; --j) Dm[j * N + i] = Dm[j * N + i] + 1.0 * Dm[(j + 1) * N + i];
                    ys[i] = yp[i];
                }
                __syncthreads();
                t = t_predict;
                st.v[DSB_STAT_STEPS] += 1;
                ju.step();
                has_prev_error = true; prev_error_norm = error_norm;
                n_equal_steps += 1;
                if (n_equal_steps > order) {
                    const int ord = order;
                    const double inf = dsb_from_bits(0x7ff0000000000000ULL);
                    double error_m_norm = inf, error_p_norm = inf;
                    if (ord > 1) { const double e = squared_norm(Dm + ord * N, ys) * pa.tab.error_const2[ord - 1]; error_m_norm = (0.0 < e) ? e : 0.0; }
                    if (ord < DSB_MAX_ORDER) { const double e = squared_norm(Dm + (ord + 2) * N, ys) * pa.tab.error_const2[ord + 1]; error_p_norm = (0.0 < e) ? e : 0.0; }
                    const double f0 = pi_controller_raw(error_m_norm, ord);
                    const double f1 = pi_controller_raw(error_norm, ord + 1);
                    const double f2 = pi_controller_raw(error_p_norm, ord + 2);
                    int max_index = 0;
                    double fmax = f0;
                    if (!(fmax > f1)) { max_index = 1; fmax = f1; }
                    if (!(fmax > f2)) { max_index = 2; fmax = f2; }
                    order = ord + (max_index - 1);
                    double factor = safety * fmax;
                    if (factor > pa.opt.max_timestep_growth) factor = pa.opt.max_timestep_growth;
                    if (factor < pa.opt.min_timestep_shrink) factor = pa.opt.min_timestep_shrink;
                    if (factor >= pa.opt.min_timestep_growth || factor <= pa.opt.max_timestep_shrink || max_index != 1) {
                        const int e = update_step_size(factor, &new_h);
                        if (e) { status = e; break; }
                        jacobian_updates(new_h * pa.tab.alpha[order], DSB_STEP_SUCCESS);
                    }
                }
                if constexpr (NR > 0) {   // also in the step()/interpolate() loop of the reference's harness (free_running), which returns interpolate(t_root) and ends (ode_solver/mod.rs:134-141)
                    // check for a root within the accepted step (bdf.rs:1566-1579), after the step-size update and before
                    // the stop time is handled; the interpolated state of the secant iteration goes to the (free) vector yc
                    double t_root = t;
                    __syncthreads();
                    const bool stopped_on_root = rf.check_root(t, [&](double (&gv)[NR]) { M::root(ys, p, t, gv); },
                                                               [&](double t_mid, double (&gv)[NR]) {
                                                                   interpolate_to_shared(t_mid, yc);
                                                                   M::root(yc, p, t_mid, gv);
                                                               }, t_root, root_found);
                    if (stopped_on_root) {
                        // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root)
                        // (bdf.rs:1228-1262), then -- without a reset function -- the state at the root in the next column
                        // (method.rs:493-503) and the end of the solve
                        while (!free_running && col < nt && bb.t_eval[col] <= t_root) {
                            (void)interpolate_and_write(bb.t_eval[col], col);
                            ++col;
                        }
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (state.rs:246-270: y <- reset(y, t),
                                // dy <- f(y, t)), a new stop time, then Bdf::step finds the state modified
                                // (bdf.rs:1291-1318): root finder re-initialised, difference array back to first order,
                                // _jacobian_updates(c, StepSuccess), set_stop_time again.  Every thread evaluates the
                                // (small) whole-vector reset and rhs functions on the shared state at the root.
                                ended = false;
                                interpolate_to_shared(t_root, yc);
                                t = t_root;
                                apply_reset_at(yc, t);
                                root_found = -1;
                                if (t < bb.t_eval[nt - 1]) {
                                    step_result = 3;
                                    has_tstop = true; tstop = bb.t_eval[nt - 1];
                                    int r = handle_tstop(tstop);                              // method.rs:792
                                    if (r == 0) {
                                        M::root(ys, p, t, rf.g0); rf.t0 = t;
                                        order = 1; n_equal_steps = 0;
                                        __syncthreads();
                                        for (int i = tid; i < N; i += T) { Dm[i] = ys[i]; Dm[N + i] = dys[i] * h; }
                                        __syncthreads();
                                        c = h * pa.tab.alpha[1];
                                        jacobian_updates(c, DSB_STEP_SUCCESS);
                                        has_prev_error = false;
                                        has_tstop = true;
                                        r = handle_tstop(tstop);                              // bdf.rs:1314-1316
                                    }
                                    if (r == 1) { has_tstop = false; status = DSB_STATUS_STOP_TIME_AT_CURRENT; break; }
                                    else if (r < 0) { status = -r; break; }
                                } else {
                                    step_result = 1;                                          // TstopReached
                                }
                            }
                        }
                        if (ended) {
                            if (col < nt) {
                                (void)interpolate_and_write(t_root, col);
                                ++col;
                            }
                            __syncthreads();
                            if (!free_running) t = t_root;      // state_mut_back; the harness loop leaves the state at the end of the step
                            step_result = 2;
                        }
                    }
                }
                if (step_result == 3) step_result = 0;             // a reset was applied: the stop time is set already
                else if (has_tstop && step_result == 0) {
                    const int r = handle_tstop(tstop);
                    if (r == 1) step_result = 1;
                    else if (r < 0) { status = -r; break; }
                }
            }
            if (step_result == 2) break;                           // RootFound ends the solve
            if (!free_running) {
                while (col < nt && bb.t_eval[col] <= t) {
                    const int e = interpolate_and_write(bb.t_eval[col], col);
                    if (e) { status = e; break; }
                    ++col;
                }
                if (step_result == 1) break;
            }
        }
        write_results(status, t, h, order, col, root_found);
        }
    }
}
