// dsb_band_bdf_kernel.cuh -- `problem.bdf::<LS>()?.solve_dense(t_eval)` for BANDED systems of medium size
// (n > 16, identity mass, df/dy inside a declared band kl, ku <= 2): method-of-lines models such as the
// single-particle battery model of BASELINE config 5 (two radial diffusion grids, tridiagonal).
//
// Execution model: still ONE LANE PER INSTANCE with the per-lane state machine and warp-level block scheduler of
// dsb_bdf_kernel.cuh, but the instance's vectors and matrices do not fit on chip any more (n = 42: 7 KB per
// instance), so they live in GLOBAL memory, one column per resident lane, word w of lane g at ws[w * LS + g]
// (LS = lanes of the whole grid).  Every vector operation is a loop over the components with the SAME trip count
// in all lanes, so the 32 lanes of a warp touch 32 consecutive doubles in every load and store: 256-byte
// coalesced accesses, the batch-major layout SURVEY.md section 8d asks for.  This kernel is bound by HBM
// bandwidth, not by FP64 issue: what it moves per Newton iteration is the band factors (4n words), the Newton
// work vectors (6n words) and nothing else.
//
// The iteration matrix I - c J is factored per lane, sequentially, in LAPACK band storage (dgbtf2 convention,
// 2 kl + ku + 1 rows per column: room for the fill-in of partial pivoting); the substitutions keep the running
// entries in a REGISTER WINDOW (kl + 1 resp. kl + ku + 1 values), so that the recurrence never waits for a
// global-memory round trip.  Same arithmetic as nalgebra's dense LU / solve (first maximum as pivot,
// reciprocal-pivot scaling, `a = (-u) * l + a` in ascending pivot order, column-axpy substitutions): every
// operation that is skipped has an exactly zero multiplier or pivot-row entry, the interchanges are interleaved
// with the forward substitution as in dsb_coop.cuh:warp_band_solve, and the solutions are bit-identical to the
// dense path (tests/test_gpu_band_parity.py against the oracle's dense LU).
//
// Restated functions: the same list as dsb_bdf_kernel.cuh, plus new_without_initialise / set_step_size
// (ode_solver/state.rs:1086-1124, 1209-1277), which the small-n path runs in dsb_init_kernel.cuh.
#pragma once
#include "dsb_band_lu.cuh"
#include "dsb_roots.cuh"
#include "dsb_bdf_kernel.cuh"

#ifndef DSB_BAND_THREADS
#define DSB_BAND_THREADS 768        // 24 warps at 80 registers: 1.16x over 16 warps at 128 (latency bound on global loads); 32 warps spill too much
#endif
#define DSB_BAND_THREADS_SMALL 128  // batches that do not fill one 768-lane block per SM are spread over the SMs in small blocks
#ifndef DSB_BAND_UNROLL_SMALL
#define DSB_BAND_UNROLL_SMALL 4     // vector loops of the small-block variant: few resident warps, so the memory-level
#endif                              // parallelism has to come from independent loads of ONE lane (255 registers available)
// block sizes of the component loops (band_for; U2: loops that evaluate the equations or carry several results per
// component, U4: plain vector loops)
#ifndef DSB_BAND_UNROLL_BIG
#define DSB_BAND_UNROLL_BIG 1     // 24 warps per SM at 80 registers: the warps provide the parallelism, bigger blocks spill
#endif
template <int T> struct BandUnroll {
    static constexpr int U2 = T <= DSB_BAND_THREADS_SMALL ? DSB_BAND_UNROLL_SMALL : DSB_BAND_UNROLL_BIG;
    static constexpr int U4 = 2 * U2;
    static constexpr int UN = U4 > 4 ? U4 : 4;         // loops without stores (norms): plain unrolling is enough
};

template <class M, int T = DSB_BAND_THREADS>
struct BandBdfLayout {
    static constexpr int N = M::N, NP = M::NP;
    static constexpr int KL = M::BAND_KL, KU = M::BAND_KU, KV = KL + KU;
    static constexpr int LDJ = KL + KU + 1;                         // rows of the band storage of df/dy (and of M)
    static constexpr int LDAB = 2 * KL + KU + 1;                    // rows of the band storage of the factors
    static constexpr int O_D = 0;                                   // D[DSB_NDIFF][N]
    static constexpr int O_Y = O_D + DSB_NDIFF * N;                 // state.y
    static constexpr int O_YP = O_Y + N;                            // y_predict
    static constexpr int O_YC = O_YP + N;                           // Newton iterate
    static constexpr int O_PSI = O_YC + N;                          // psi - y_predict
    static constexpr int O_DL = O_PSI + N;                          // Newton residual / update
    static constexpr int O_J = O_DL + N;                            // df/dy, band storage: (i, j) at j * LDJ + KU + i - j
    static constexpr int O_LU = O_J + LDJ * N;                      // factors, band storage: (i, j) at j * LDAB + KV + i - j
    static constexpr int O_PIV = O_LU + LDAB * N;                   // pivot offsets (row j interchanged with row j + piv[j])
    static constexpr int O_RU = O_PIV + N;                          // rescale matrix R U when it does not fit shared memory
    static constexpr int O_M = O_RU + 25;                           // mass matrix, band storage like df/dy (DAEs only)
    static constexpr int O_TMP = O_M + (M::HAS_MASS ? LDJ * N : 0); // y + psi - y_predict, the argument of M (DAEs only)
    static constexpr int WORDS = O_TMP + (M::HAS_MASS ? N : 0);
    static constexpr int THREADS = T;
    static constexpr int MAXNREG = (65536 / THREADS) / 8 * 8 > 255 ? 255 : (65536 / THREADS) / 8 * 8;
    static constexpr bool RU_IN_SMEM = THREADS <= 768;
    static constexpr int SMEM_WORDS = (DSB_NSTATS + 1) / 2 + (RU_IN_SMEM ? 25 : 0);   // statistics (+ rows / columns 1..5 of R U)
    static_assert(KL >= 1 && KL <= 2 && KU >= 1 && KU <= 2, "register windows are sized for kl, ku <= 2");
};

// indexable view of one vector of the lane's global-memory column (what the component-wise equations read)
struct BandVec {
    const double* base; size_t ls;
    __device__ __forceinline__ double operator[](int k) const { return base[(size_t)k * ls]; }
};
// Per-column metadata of df/dy (device array, the same for every instance: jacobian/mod.rs:32): bits 0-15 the colour of
// the column, bits 16.. the sparsity pattern of the column inside the band (bit 16 + KU + i - j <=> entry (i, j)).
struct DsbBandMeta {
    const double* atol;        // [n]
    const int32_t* colmeta;    // [n]
};
// seed of one colour: 1 in every column of the colour that has a non-zero
struct BandColourSeed {
    const int32_t* colmeta; int c;
    __device__ __forceinline__ double operator[](int k) const {
        const int32_t m = colmeta[k];
        return ((m & 0xffff) == c && (m >> 16) != 0) ? 1.0 : 0.0;
    }
};

// per-component results of the blocked loops (dsb_band_lu.cuh: band_for)
struct BandR2 { double a, b; };
struct BandRCols { double v[DSB_MAX_ORDER + 1]; };
struct BandRDiff { double d2, d1, yp, col[DSB_MAX_ORDER + 1]; };
template <int LD> struct BandRBand { double v[LD]; };

// e_j: the argument of `mass` when the mass matrix is assembled column by column (op/linear_op.rs:42-51)
struct BandUnitVec {
    int j;
    __device__ __forceinline__ double operator[](int k) const { return k == j ? 1.0 : 0.0; }
};

template <class M, int T>
__global__ void __maxnreg__((BandBdfLayout<M, T>::MAXNREG)) dsb_band_bdf_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                                  const __grid_constant__ DsbBatchBuffers bb,
                                                                  const __grid_constant__ DsbBandMeta meta,
                                                                  double* __restrict__ ws,
                                                                  unsigned long long* __restrict__ work_counter) {
    typedef BandBdfLayout<M, T> Lay;
    constexpr int U2 = BandUnroll<T>::U2, U4 = BandUnroll<T>::U4, UN = BandUnroll<T>::UN;
    typedef LaneBandLU<M::N, Lay::KL, Lay::KU, DsbDivShared, U2> BLU;
    constexpr int N = Lay::N, NP = Lay::NP, KL = Lay::KL, KU = Lay::KU, KV = Lay::KV, LDJ = Lay::LDJ, LDAB = Lay::LDAB;
    extern __shared__ double dsb_lane_smem[];
    double* const sm = dsb_lane_smem + threadIdx.x;
#define SMW(w) sm[(w) * Lay::THREADS]
#define SRU(i, j) (*(Lay::RU_IN_SMEM ? &SMW((DSB_NSTATS + 1) / 2 + ((i) - 1) * 5 + ((j) - 1)) : &g[(size_t)(Lay::O_RU + ((i) - 1) * 5 + ((j) - 1)) * LS]))
    const size_t LS = (size_t)gridDim.x * blockDim.x;
    double* const g = ws + ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
#define G(w) g[(size_t)(w) * LS]
#define GD(j, i) G(Lay::O_D + (j) * N + (i))
#define GY(i) G(Lay::O_Y + (i))
#define GYP(i) G(Lay::O_YP + (i))
#define GYC(i) G(Lay::O_YC + (i))
#define GPSI(i) G(Lay::O_PSI + (i))
#define GDL(i) G(Lay::O_DL + (i))
#define GJ(j, r) G(Lay::O_J + (j) * LDJ + (r))
#define GAB(j, r) G(Lay::O_LU + (j) * LDAB + (r))
#define GPIV(j) G(Lay::O_PIV + (j))
#define GM(j, r) G(Lay::O_M + (j) * LDJ + (r))
#define GTMP(i) G(Lay::O_TMP + (i))
#define DSB_DIV(a, b) DsbDivShared::div((a), (b))
    const BandVec vY{g + (size_t)Lay::O_Y * LS, LS}, vYC{g + (size_t)Lay::O_YC * LS, LS}, vTMP{g + (size_t)Lay::O_TMP * LS, LS},
                  vDL{g + (size_t)Lay::O_DL * LS, LS};

    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const int quorum = pa.quorum;
    const int newton_passes = pa.newton_passes < 1 ? 1 : pa.newton_passes;       // see dsb_bdf_kernel.cuh (NEWTON block)
    const double eps = 2.220446049250313e-16;

    // ---- per-lane registers (the controller of dsb_bdf_kernel.cuh) ----------------------------------------
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
    SmemLaneStats<2 * Lay::THREADS> st;
    st.v.base = reinterpret_cast<int*>(&SMW(0));
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
    auto finish = [&](int status) { fin_status = status; state = L_FINISH; };
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
    // ||x||^2_w(ref) (vector/nalgebra_serial.rs:395-408): x and ref are word offsets of the lane's column; the
    // terms are added in index order, the loads do not depend on the sum and run ahead of it
    auto weighted_norm = [&](int ox, int oref) -> double {
        double acc = 0.0;
#pragma unroll UN
        for (int i = 0; i < N; ++i) {
            const double term = DSB_DIV(G(ox + i), dsb_abs(G(oref + i)) * pa.rtol + meta.atol[i]);
            acc += term * term;
        }
        return DSB_DIV(acc, (double)N);
    };

    // root finding (dsb_roots.cuh; bdf.rs:143, 301-306, 1566-1579): only compiled for equations with roots
    constexpr int NR = dsb_model_nroots<M>::value;
    LaneRootFinder<(NR > 0 ? NR : 1), DsbDivShared> rf;
    rf.t0 = 0.0;
    int root_found = -1;
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
    // interpolate (bdf.rs:767-782, 1080-1106): the time factors first, then one pass over the components
    auto interpolate_to = [&](double tq, auto&& store) {
        double tf[DSB_MAX_ORDER];
        double time_factor = 1.0;
#pragma unroll
        for (int j = 0; j < DSB_MAX_ORDER; ++j) {
            if (j < order) {
                const double j_t = (double)j;
                time_factor *= DSB_DIV(tq - (t - h * j_t), h * (1.0 + j_t));
            }
            tf[j] = time_factor;
        }
        band_for<U2, double>(N, [&](int i) {
            double yo = GD(0, i);
#pragma unroll
            for (int j = 0; j < DSB_MAX_ORDER; ++j) if (j < order) yo = tf[j] * GD(j + 1, i) + yo;
            return yo;
        }, store);
    };
    // the same into the (free) Newton residual vector, for the output and root functions: only the components they
    // read when the equations declare them (dsb_math.h: dsb_model_ndep)
    constexpr int NDEP = dsb_model_ndep<M>::value;
    auto interpolate_for_functions = [&](double tq) {
        if constexpr (NDEP > 0) {
            double tf[DSB_MAX_ORDER];
            double time_factor = 1.0;
#pragma unroll
            for (int j = 0; j < DSB_MAX_ORDER; ++j) {
                if (j < order) {
                    const double j_t = (double)j;
                    time_factor *= DSB_DIV(tq - (t - h * j_t), h * (1.0 + j_t));
                }
                tf[j] = time_factor;
            }
            band_for<NDEP, double>(NDEP, [&](int q) {
                const int i = M::dep(q);
                double yo = GD(0, i);
#pragma unroll
                for (int j = 0; j < DSB_MAX_ORDER; ++j) if (j < order) yo = tf[j] * GD(j + 1, i) + yo;
                return yo;
            }, [&](int q, double yo) { GDL(M::dep(q)) = yo; });
        } else {
            interpolate_to(tq, [&](int i, double yo) { GDL(i) = yo; });
        }
    };

    // one column of the solve_dense result (dense_write_out, method.rs:822-848): the interpolated state, or -- for
    // equations with an output function -- out(y(tq), tq), evaluated on the state interpolated into the (free) Newton
    // residual vector
    constexpr int NOUT = dsb_model_nout<M>::value;
    auto write_column = [&](double tq, int column) {
        if constexpr (dsb_model_nout<M>::has_out) {
            interpolate_for_functions(tq);
            double o[NOUT];
            M::out(vDL, pl, tq, o);
#pragma unroll
            for (int k = 0; k < NOUT; ++k) bb.ys[((int64_t)column * NOUT + k) * B + inst] = o[k];
        } else {
            interpolate_to(tq, [&](int i, double yo) { bb.ys[((int64_t)column * N + i) * B + inst] = yo; });
        }
    };

    while (true) {
        // ---- warp-level block scheduler (dsb_bdf_kernel.cuh) ---------------------------------------------------
        const unsigned m_idle = __ballot_sync(0xffffffffu, state == L_IDLE);
        if (m_idle == 0xffffffffu) break;
        const int n_active = 32 - __popc(m_idle);
        const int n_slow = __popc(__ballot_sync(0xffffffffu, state == L_SELECT || state == L_RESCALE || state == L_JAC));
        const bool run_slow = n_slow > 0 && (n_slow >= quorum || 2 * n_slow >= n_active);

        // ================= FINISH =================================================================================
        if (__any_sync(0xffffffffu, state == L_FINISH) && state == L_FINISH) {
            bb.status[inst] = fin_status;
            bb.fin_t[inst] = t; bb.fin_h[inst] = h; bb.fin_order[inst] = order;
#pragma unroll
            for (int k = 0; k < DSB_NSTATS; ++k) bb.stats[(int64_t)k * B + inst] = st.v[k];
            if (NR > 0) { bb.ncols[inst] = col; bb.root_idx[inst] = root_found; }
            state = L_FETCH;
        }
        // ================= FETCH: next instance; new_without_initialise, set_step_size, Bdf::_new part 1 ==============
        if (__any_sync(0xffffffffu, state == L_FETCH) && state == L_FETCH) {
            inst = (int64_t)atomicAdd(work_counter, 1ull);
            if (inst >= B) {
                state = L_IDLE;
            } else if (!M::HAS_MASS || bb.status[inst] == DSB_STATUS_OK) {     // else: consistent initialisation failed, keep its status
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = bb.params[(int64_t)j * B + inst];
#pragma unroll
                for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = 0;
                t = pa.t0;
                if constexpr (M::HAS_MASS) {
                    // singular mass: y, dy after set_consistent and the counters so far come from dsb_band_init_kernel
                    // (state.rs:84-162)
#pragma unroll
                    for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = bb.stats[(int64_t)k * B + inst];
                    band_for<U4, BandR2>(N, [&](int i) { return BandR2{bb.y0[(int64_t)i * B + inst], bb.dy0[(int64_t)i * B + inst]}; },
                                         [&](int i, const BandR2& r) { GY(i) = r.a; GD(1, i) = r.b; });
                } else {
                    // y = init(p, t0); dy = f(y, t0)     (state.rs:1086-1124); dy is kept in D[1] until h is known
#pragma unroll UN
                    for (int i = 0; i < N; ++i) GY(i) = M::init_i(i, pl, pa.t0);
                    band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vY, pl, pa.t0); }, [&](int i, double r) { GD(1, i) = r; });
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                }
                // set_step_size (state.rs:1209-1277), solver order 1
                {
                    const bool is_neg_h = pa.h0 < 0.0;
                    const double d0 = dsb_sqrt(weighted_norm(Lay::O_Y, Lay::O_Y));
                    const double d1 = dsb_sqrt(weighted_norm(Lay::O_D + N, Lay::O_Y));
                    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * DSB_DIV(d0, d1);
                    band_for<U4, double>(N, [&](int i) { return is_neg_h ? (GD(1, i) * (-h0) + GY(i)) : (GD(1, i) * h0 + GY(i)); },
                                         [&](int i, double r) { GYC(i) = r; });
                    const double t1 = is_neg_h ? pa.t0 - h0 : pa.t0 + h0;
                    band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vYC, pl, t1) - GD(1, i); }, [&](int i, double r) { GDL(i) = r; });
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                    const double d2 = DSB_DIV(dsb_sqrt(weighted_norm(Lay::O_DL, Lay::O_Y)), dsb_abs(h0));
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
                band_for<U4, BandR2>(N, [&](int i) { return BandR2{GY(i), GD(1, i) * h}; },
                                     [&](int i, const BandR2& r) {
                                         GD(0, i) = r.a; GD(1, i) = r.b;
#pragma unroll
                                         for (int j = 2; j < DSB_NDIFF; ++j) GD(j, i) = 0.0;
                                     });
                order = 1; n_equal_steps = 0;
                conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
                c = h * pa.tab.alpha[1];
                jacobian_is_stale = true;
                ju.init(1.0);                                   // jacobian_update.rs:27 -- h_at_last starts at ONE
                has_tstop = false; tstop = 0.0; has_prev_error = false; prev_error_norm = 0.0;
                convergence_fail = false; first = true; reached = false; pending_etf = false; col = 0;
                t_predict = t;
                if constexpr (NR > 0) {                         // Bdf::_new: root_finder.init(root_fn, state.y, state.t)
                    M::root(vY, pl, t, rf.g0);
                    rf.t0 = t; root_found = -1;
                }
                jac_kind = DSB_KIND_CONSTRUCT;
                state = L_JAC;
            }
        }
        // ================= REINIT: Bdf::step finds the state modified by a reset (bdf.rs:1291-1318) =========================
        // root finder re-initialised, difference array back to first order (D[:, 0] = y, D[:, 1] = h dy; the higher columns
        // keep what they held), _jacobian_updates(c, StepSuccess), then set_stop_time again: the TSTOP block's first-step path
        if constexpr (dsb_model_has_reset<M>::value) {
            if (__any_sync(0xffffffffu, state == L_REINIT) && state == L_REINIT) {
                M::root(vY, pl, t, rf.g0);
                rf.t0 = t;
                order = 1; n_equal_steps = 0;
                band_for<U4, BandR2>(N, [&](int i) { return BandR2{GY(i), GYP(i) * h}; },
                                     [&](int i, const BandR2& r) { GD(0, i) = r.a; GD(1, i) = r.b; });
                c = h * pa.tab.alpha[1];
                has_prev_error = false;
                jac_kind = DSB_STEP_SUCCESS; after_jac = L_TSTOP;
                first = true;
                state = L_JAC;
            }
        }
        // ================= SELECT (bdf.rs:1489-1563, 1431-1442) ===========================================================
        if (run_slow && state == L_SELECT) {
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
                            const double e = weighted_norm(Lay::O_D + (ord + q) * N, Lay::O_Y) * pa.tab.error_const2[ord - 1 + q];
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
        // R U (rows / columns 1..k; row and column 0 are those of the identity) is built row by row as in
        // dsb_bdf_kernel.cuh and parked in shared memory, then D[:, 1..k] <- D[:, 1..k] (R U) one component at a time.
        if (run_slow && state == L_RESCALE) {
            const double factor = rescale_factor;
            const double new_h = factor * h;
            n_equal_steps = 0;
            const int k = order;
            const double* __restrict__ u = pa.tab.u[DSB_MAX_ORDER];         // leading dimension 6
            {
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
            band_for<U2, BandRCols>(N, [&](int s) {
                BandRCols nd;
#pragma unroll
                for (int j = 1; j <= DSB_MAX_ORDER; ++j) nd.v[j] = -0.0;    // (-0.0) + x == x: the first term is assigned
#pragma unroll 1
                for (int i = 1; i <= k; ++i) {
                    const double di = GD(i, s);
#pragma unroll
                    for (int j = 1; j <= DSB_MAX_ORDER; ++j) nd.v[j] = di * SRU(i, j) + nd.v[j];
                }
                return nd;
            }, [&](int s, const BandRCols& nd) {
#pragma unroll
                for (int j = 1; j <= DSB_MAX_ORDER; ++j) if (j <= k) GD(j, s) = nd.v[j];
            });
            c = new_h * pa.tab.alpha[k];
            h = new_h;
            conv.eta = pa.tab.eta_reset_timestep;
            if (!rs_ignore_small && dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
            else state = after_rescale;
        }

        // ================= JAC: _jacobian_updates(c, kind) / Bdf::_new's reset_jacobian =====================================
        if (run_slow && state == L_JAC) {
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
                    // the sparsity pattern into band storage
                    st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
                    for (int e = 0; e < LDJ * N; ++e) G(Lay::O_J + e) = 0.0;
                    const bool one_colour_per_column = pa.ncolors == N;
#pragma unroll 1
                    for (int cc = 0; cc < pa.ncolors; ++cc) {
                        const BandColourSeed seed{meta.colmeta, cc};
                        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
                        // rows outside the band of the colour's only column hold exact zeros and are not evaluated
                        const int i0 = one_colour_per_column ? (cc - KU < 0 ? 0 : cc - KU) : 0;
                        const int i1 = one_colour_per_column ? (cc + KL > N - 1 ? N - 1 : cc + KL) : N - 1;
                        band_for<U2, double>(i1 - i0 + 1, [&](int q) { return M::jac_mul_i(i0 + q, vY, pl, t, seed); },
                                             [&](int q, double val) {
                            const int i = i0 + q;
#pragma unroll
                            for (int d = -KL; d <= KU; ++d) {               // column j = i + d
                                const int j = i + d;
                                if (j >= 0 && j < N) {
                                    const int32_t m = meta.colmeta[j];
                                    if ((m & 0xffff) == cc && ((m >> (16 + KU - d)) & 1)) GJ(j, KU - d) = val;
                                }
                            }
                        });
                    }
                    if constexpr (M::HAS_MASS) {
                        // mass.matrix_inplace(t) with the Jacobian (op/bdf.rs:273-300): column j = M e_j, beta = 0
#pragma unroll 1
                        for (int j = 0; j < N; ++j) {
                            const BandUnitVec ej{j};
#pragma unroll
                            for (int r = 0; r < LDJ; ++r) {
                                const int i = j + r - KU;
                                GM(j, r) = (i >= 0 && i < N) ? M::mass_i(i, ej, pl, t, 0.0, 0.0) : 0.0;
                            }
                        }
                    }
                    jacobian_is_stale = false;
                }
                // A = M - c J (op/bdf.rs:282-298: J * (-c) + M) in band storage with kl extra rows for the fill-in
                const double mc = -c;
                band_for<U2, BandRBand<LDAB>>(N, [&](int j) {
                    BandRBand<LDAB> a;
#pragma unroll
                    for (int r = 0; r < LDAB; ++r) {
                        const int i = j + r - KV;
                        double v = 0.0;
                        if (r >= KL && i >= 0 && i < N) {
                            if constexpr (M::HAS_MASS) v = GJ(j, r - KL) * mc + GM(j, r - KL);
                            else v = GJ(j, r - KL) * mc + ((i == j) ? 1.0 : 0.0);
                        }
                        a.v[r] = v;
                    }
                    return a;
                }, [&](int j, const BandRBand<LDAB>& a) {
#pragma unroll
                    for (int r = 0; r < LDAB; ++r) GAB(j, r) = a.v[r];
                });
                // band LU, dgbtf2 convention (dsb_band_lu.cuh)
                BLU::factor(g, LS, Lay::O_LU, Lay::O_PIV);
            }
            state = after_jac;
        }

        // ================= TSTOP ==========================================================================================
        if (__any_sync(0xffffffffu, state == L_TSTOP) && state == L_TSTOP) {
            bool stopped_on_root = false;
            bool reset_now = false;             // a reset was applied at a root: set_stop_time again, then L_REINIT
            if constexpr (NR > 0) {
                // check for a root within the accepted step (bdf.rs:1566-1579), after the step-size update and before the
                // stop time is handled; the interpolated state of the secant iteration goes to the (free) Newton residual
                if (!first) {   // also in the step()/interpolate() loop of the reference's harness (free_running), which returns interpolate(t_root) and ends (ode_solver/mod.rs:134-141)
                    double t_root = t;
                    stopped_on_root = rf.check_root(t, [&](double (&gv)[NR]) { M::root(vY, pl, t, gv); },
                                                    [&](double t_mid, double (&gv)[NR]) {
                                                        interpolate_for_functions(t_mid);
                                                        M::root(vDL, pl, t_mid, gv);
                                                    }, t_root, root_found);
                    if (stopped_on_root) {
                        // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root)
                        // (bdf.rs:1228-1262), then -- without a reset function -- the state at the root in the next column
                        // (method.rs:493-503) and the end of the solve
                        while (!free_running && col < nt && bb.t_eval[col] <= t_root) {
                            write_column(bb.t_eval[col], col);
                            ++col;
                        }
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (state.rs:246-270: y <- reset(y, t),
                                // dy <- f(y, t); dy is parked in the predictor's vector until the difference array is
                                // re-initialised), then a new stop time and on with the integration -- or TstopReached
                                interpolate_to(t_root, [&](int i, double yo) { GDL(i) = yo; });
                                t = t_root;
                                band_for<U2, double>(N, [&](int i) { return M::reset_i(i, vDL, pl, t); }, [&](int i, double r) { GY(i) = r; });
                                band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vY, pl, t); }, [&](int i, double r) { GYP(i) = r; });
                                st.v[DSB_STAT_RHS_CALLS] += 1;
                                root_found = -1;
                                ended = false;
                                if (t < bb.t_eval[nt - 1]) { reset_now = true; stopped_on_root = false; }
                                else finish(DSB_STATUS_OK);                        // TstopReached
                            }
                        }
                        if (ended) {
                            if (col < nt) {
                                write_column(t_root, col);
                                ++col;
                            }
                            if (!free_running) t = t_root;      // state_mut_back; the harness loop leaves the state at the end of the step
                            finish(DSB_STATUS_OK);
                        }
                    }
                }
            }
            int next = first ? L_PREDICT : L_OUTPUT;
            int r = 0;
            bool check = has_tstop && !stopped_on_root;
            if (reset_now) { next = L_REINIT; check = true; has_tstop = true; tstop = bb.t_eval[nt - 1]; }
            if (first) {
                check = !free_running;
                if (free_running) next = L_OUTPUT;
                else { has_tstop = true; tstop = bb.t_eval[nt - 1]; }
            }
            if (check) {
                r = handle_tstop(tstop);
                if (r == 1) {
                    if (first || reset_now) r = -DSB_STATUS_STOP_TIME_AT_CURRENT;
                    else reached = true;
                }
            }
            if (stopped_on_root) {
                // the lane is on its way to FINISH
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
        if (__any_sync(0xffffffffu, state == L_OUTPUT) && state == L_OUTPUT) {
            int status = DSB_STATUS_OK;
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
        if (__any_sync(0xffffffffu, state == L_PREDICT) && state == L_PREDICT) {
            if (repredict) {
                const int ord = order;
                const double a = pa.tab.alpha[ord];
                band_for<U2, BandR2>(N, [&](int i) {
                    double yp = 0.0;
                    double ps = 0.0;
#pragma unroll
                    for (int j = 0; j <= DSB_MAX_ORDER; ++j) {
                        if (j <= ord) {
                            const double d = GD(j, i);
                            yp += d;
                            if (j == 1) ps = pa.tab.gamma[1] * d;
                            else if (j >= 2) ps = pa.tab.gamma[j] * d + ps;
                        }
                    }
                    ps *= a;
                    ps -= yp;
                    return BandR2{yp, ps};
                }, [&](int i, const BandR2& r) { GYP(i) = r.a; GPSI(i) = r.b; GYC(i) = r.a; });
                t_predict = t + h;
            } else {
                band_for<U4, double>(N, [&](int i) { return GYP(i); }, [&](int i, double r) { GYC(i) = r; });
            }
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
#pragma unroll 1
        for (int pass = 0; pass < newton_passes; ++pass) {
        if (!__any_sync(0xffffffffu, state == L_NEWTON)) break;
        if (state == L_NEWTON) {
            // delta = F(y) = M (y + psi - y0) - c f(t, y)   (op/bdf.rs:240-256)
            const double mc = -c;
            if constexpr (M::HAS_MASS) {
                band_for<U4, double>(N, [&](int i) { return GYC(i) + GPSI(i); }, [&](int i, double r) { GTMP(i) = r; });
                band_for<U2, double>(N, [&](int i) {
                    const double f = M::rhs_i(i, vYC, pl, t_predict);
                    return M::mass_i(i, vTMP, pl, t_predict, mc, f);         // gemv_inplace(x, t, beta, y): y = M x + beta y
                }, [&](int i, double r) { GDL(i) = r; });
            } else {
                band_for<U2, double>(N, [&](int i) {
                    const double f = M::rhs_i(i, vYC, pl, t_predict);
                    return (GYC(i) + GPSI(i)) + mc * f;
                }, [&](int i, double r) { GDL(i) = r; });
            }
            st.v[DSB_STAT_RHS_CALLS] += 1;
            const bool ok = BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, Lay::O_DL);
            if (!ok) {
                newton_ok = false; state = L_POST;              // LuSolveFailed
            } else {
                double acc = 0.0;
                band_for<U4, BandR2>(N, [&](int i) {
                    const double dl = GDL(i);
                    // Newton norm weights use the PREDICTOR (line_search.rs:67, convergence.rs:64-66)
                    return BandR2{GYC(i) - dl, DSB_DIV(dl, dsb_abs(GYP(i)) * pa.rtol + meta.atol[i])};
                }, [&](int i, const BandR2& r) { GYC(i) = r.a; acc += r.b * r.b; });
                const double norm = dsb_sqrt(DSB_DIV(acc, (double)N));
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
        }
        // ================= POST: a Newton solve ended (bdf.rs:1338-1563) =====================================================
        if (__any_sync(0xffffffffu, state == L_POST) && state == L_POST) {
            st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
            if (newton_ok) {
                const int ord = order;
                {   // error_control: ||d||^2_w(state.y) * error_const2[order - 1], d = y - y_predict
                    double acc = 0.0;
#pragma unroll UN
                    for (int i = 0; i < N; ++i) {
                        const double d = GYC(i) - GYP(i);
                        const double term = DSB_DIV(d, dsb_abs(GY(i)) * pa.rtol + meta.atol[i]);
                        acc += term * term;
                    }
                    const double err = DSB_DIV(acc, (double)N) * pa.tab.error_const2[ord - 1];
                    error_norm = (0.0 < err) ? err : 0.0;
                }
                const double maxiter = (double)conv.max_iter;
                const double niter = (double)conv.niter;
                safety = DSB_DIV(0.9 * (2.0 * maxiter + 1.0), 2.0 * maxiter + niter);
                if (error_norm <= 1.0) {
                    // ---- accepted: _update_diff, state.y <- PREDICTOR (quirk Q1) ----
                    band_for<(U2 > 4 ? 4 : U2), BandRDiff>(N, [&](int i) {
                        BandRDiff r;
                        r.yp = GYP(i);
                        const double d = GYC(i) - r.yp;
                        double above = d;                                   // the new D[:, ord + 1]
                        r.d2 = d - GD(ord + 1, i);
                        r.d1 = d;
#pragma unroll
                        for (int j = DSB_MAX_ORDER; j >= 0; --j) {
                            if (j <= ord) {
                                above = GD(j, i) + 1.0 * above;
                                r.col[j] = above;
                            }
                        }
                        return r;
                    }, [&](int i, const BandRDiff& r) {
                        GD(ord + 2, i) = r.d2;
                        GD(ord + 1, i) = r.d1;
#pragma unroll
                        for (int j = DSB_MAX_ORDER; j >= 0; --j) if (j <= ord) GD(j, i) = r.col[j];
                        GY(i) = r.yp;
                    });
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
#undef SMW
#undef SRU
#undef G
#undef GD
#undef GY
#undef GYP
#undef GYC
#undef GPSI
#undef GDL
#undef GJ
#undef GAB
#undef GPIV
#undef GM
#undef GTMP
#undef DSB_DIV
}
