// dsb_band_sdirk_kernel.cuh -- `problem.tr_bdf2::<LS>()?.solve_dense(t_eval)` / `esdirk34` for BANDED systems of
// medium size (n > 16, df/dy and M inside a declared band kl, ku <= 2): the (E)SDIRK state machine of
// dsb_sdirk_kernel.cuh on the execution model of dsb_band_bdf_kernel.cuh -- one lane per instance, the instance's
// vectors and matrices in the lane's global-memory column (word w of lane g at ws[w * LS + g]: every vector loop
// is a coalesced 256-byte access per warp), M - (gamma h) J factored per lane in band storage (dsb_band_lu.cuh).
//
// Restated functions: the list of dsb_sdirk_kernel.cuh (Sdirk::step ode_solver/sdirk.rs:409-543, the Rk core
// ode_solver/runge_kutta.rs:446-960, SdirkCallable op/sdirk.rs:157-292, newton_iteration + Convergence), plus
// new_without_initialise / set_step_size (ode_solver/state.rs:1086-1124, 1209-1277) in the FETCH block; singular
// mass matrices get their consistent initial state from dsb_band_init_kernel.cuh.  All the reference quirks listed
// in dsb_sdirk_kernel.cuh are kept; vector operations whose dense form adds exactly-zero out-of-band products skip them.
#pragma once
#include "dsb_band_bdf_kernel.cuh"
#include "dsb_sdirk_kernel.cuh"

template <class M, int T = DSB_BAND_THREADS>
struct BandSdirkLayout {
    static constexpr int N = M::N, NP = M::NP;
    static constexpr int KL = M::BAND_KL, KU = M::BAND_KU, KV = KL + KU;
    static constexpr int LDJ = KL + KU + 1, LDAB = 2 * KL + KU + 1;
    // the first words coincide with BandBdfLayout where dsb_band_init_kernel.cuh needs them (it only uses words below
    // BandBdfLayout::O_RU and O_M; both layouts are sized by the larger of the two)
    static constexpr int O_DIFF = 0;                                // diff[DSB_RK_MAX_STAGES][N]: x_i = h k_i
    static constexpr int O_Y = O_DIFF + DSB_RK_MAX_STAGES * N;      // state.y
    static constexpr int O_DY = O_Y + N;                            // state.dy
    static constexpr int O_OY = O_DY + N;                           // old_state.y (last stage value / previous step)
    static constexpr int O_PHI = O_OY + N;                          // SdirkCallable.phi
    static constexpr int O_X = O_PHI + N;                           // Newton iterate (old_state.dy)
    static constexpr int O_DL = O_X + N;                            // Newton residual / update, error estimate
    static constexpr int O_TMP = O_DL + N;                          // phi + c x (argument of f), M-product argument
    static constexpr int O_J = O_TMP + N;                           // df/dy, band storage
    static constexpr int O_LU = O_J + LDJ * N;                      // factors, band storage
    static constexpr int O_PIV = O_LU + LDAB * N;                   // pivot offsets
    static constexpr int O_M = O_PIV + N;                           // mass matrix, band storage (DAEs only)
    static constexpr int WORDS_OWN = O_M + (M::HAS_MASS ? LDJ * N : 0);
    static constexpr int WORDS = WORDS_OWN > BandBdfLayout<M, T>::WORDS ? WORDS_OWN : BandBdfLayout<M, T>::WORDS;
    static constexpr int THREADS = T;
    static constexpr int MAXNREG = (65536 / THREADS) / 8 * 8 > 255 ? 255 : (65536 / THREADS) / 8 * 8;
    static constexpr int SMEM_WORDS = (DSB_NSTATS + 1) / 2;        // statistics
};

template <class M, int T>
__global__ void __maxnreg__((BandSdirkLayout<M, T>::MAXNREG)) dsb_band_sdirk_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                                    const __grid_constant__ DsbBatchBuffers bb,
                                                                    const __grid_constant__ DsbBandMeta meta,
                                                                    double* __restrict__ ws,
                                                                    unsigned long long* __restrict__ work_counter) {
    typedef BandSdirkLayout<M, T> Lay;
    constexpr int U2 = BandUnroll<T>::U2, U4 = BandUnroll<T>::U4, UN = BandUnroll<T>::UN;
    typedef LaneBandLU<M::N, Lay::KL, Lay::KU, DsbDivShared, U2> BLU;
    constexpr int N = Lay::N, NP = Lay::NP, KL = Lay::KL, KU = Lay::KU, KV = Lay::KV, LDJ = Lay::LDJ, LDAB = Lay::LDAB;
    extern __shared__ double dsb_lane_smem[];
    double* const sm = dsb_lane_smem + threadIdx.x;
#define SMW(w) sm[(w) * Lay::THREADS]
    const size_t LS = (size_t)gridDim.x * blockDim.x;
    double* const g = ws + ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
#define G(w) g[(size_t)(w) * LS]
#define GDF(j, i) G(Lay::O_DIFF + (j) * N + (i))
#define GY(i) G(Lay::O_Y + (i))
#define GDY(i) G(Lay::O_DY + (i))
#define GOY(i) G(Lay::O_OY + (i))
#define GPHI(i) G(Lay::O_PHI + (i))
#define GX(i) G(Lay::O_X + (i))
#define GDL(i) G(Lay::O_DL + (i))
#define GTMP(i) G(Lay::O_TMP + (i))
#define GJ(j, r) G(Lay::O_J + (j) * LDJ + (r))
#define GAB(j, r) G(Lay::O_LU + (j) * LDAB + (r))
#define GM(j, r) G(Lay::O_M + (j) * LDJ + (r))
#define DSB_DIV(a, b) DsbDivShared::div((a), (b))
    const BandVec vY{g + (size_t)Lay::O_Y * LS, LS}, vTMP{g + (size_t)Lay::O_TMP * LS, LS}, vX{g + (size_t)Lay::O_X * LS, LS},
                  vDL{g + (size_t)Lay::O_DL * LS, LS};

    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const int quorum = pa.quorum;
    const int newton_passes = pa.newton_passes < 1 ? 1 : pa.newton_passes;       // see dsb_bdf_kernel.cuh (NEWTON block)
    const double eps = 2.220446049250313e-16;
    const int ns = pa.rk.s;
    const int start = (pa.rk.a[0] == 0.0) ? 1 : 0;        // skip_first_stage (runge_kutta.rs:286-288)
    const double cg = pa.rk.a[1 * ns + 1];                // Sdirk::gamma() = a(1, 1)

    // ---- per-lane registers (the controller of dsb_sdirk_kernel.cuh) -----------------------------------------
    int state = R_FETCH;
    int64_t inst = 0;
    double t = 0.0, h_state = 0.0, old_t = 0.0;           // state.t, state.h, old_state.t
    double h = 0.0, op_h = 0.0;                            // step()'s local h, SdirkCallable.h
    bool has_tstop = false, has_prev_error = false, jacobian_is_stale = true, is_jacobian_set = false;
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
    int stage = 0, nattempts = 0, col = 0;
    bool updated_jacobian = false, newton_ok = false, first = true, reached = false;
    double t_stage = 0.0, factor = 1.0, error_norm = 0.0;
    int after_jac = R_NEWTON, jac_kind = DSB_CHECKPOINT;
    double jac_h = 0.0;
    int fin_status = DSB_STATUS_OK;
    auto finish = [&](int status) { fin_status = status; state = R_FINISH; };

    // runge_kutta.rs:752-781.  0 = nothing, 1 = TstopReached, < 0 = -status
    auto handle_tstop = [&](double ts) -> int {
        const double troundoff = 100.0 * eps * (dsb_abs(t) + dsb_abs(h_state));
        if (dsb_abs(t - ts) <= troundoff) return 1;
        if ((h_state > 0.0 && ts < t - troundoff) || (h_state < 0.0 && ts > t + troundoff)) return -DSB_STATUS_STOP_TIME_BEFORE_CURRENT;
        if ((h_state > 0.0 && t + h_state > ts + troundoff) || (h_state < 0.0 && t + h_state < ts - troundoff)) {
            const double f = DSB_DIV(ts - t, h_state);
            h_state *= f;
        }
        return 0;
    };
    // ||x||^2_w(ref) (vector/nalgebra_serial.rs:395-408) over words of the lane's column
    auto weighted_norm = [&](int ox, int oref) -> double {
        double acc = 0.0;
#pragma unroll UN
        for (int i = 0; i < N; ++i) {
            const double term = DSB_DIV(G(ox + i), dsb_abs(G(oref + i)) * pa.rtol + meta.atol[i]);
            acc += term * term;
        }
        return DSB_DIV(acc, (double)N);
    };

    // root finding (dsb_roots.cuh; runge_kutta.rs:43, 142-147, 935-948): only compiled for equations with roots
    constexpr int NR = dsb_model_nroots<M>::value;
    LaneRootFinder<(NR > 0 ? NR : 1), DsbDivShared> rf;
    rf.t0 = 0.0;
    int root_found = -1;
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
    // interpolate_inplace (runge_kutta.rs:1080-1127; :962-981 beta dense output, :1004-1024 Hermite) on [old_t, t]
    auto interpolate_range = [&](double tq, auto&& index, const int count, auto&& store) {
        const double dt = t - old_t;
        const double theta = (dt == 0.0) ? 1.0 : DSB_DIV(tq - old_t, dt);
        if (pa.rk.has_beta) {
            const double th2 = theta * theta;
            double bf[DSB_RK_MAX_STAGES];
#pragma unroll
            for (int j = 0; j < DSB_RK_MAX_STAGES; ++j) {
                bf[j] = 0.0;
                if (j < ns) { bf[j] = pa.rk.beta[j] * theta; bf[j] = pa.rk.beta[ns + j] * th2 + bf[j]; }
            }
            band_for<U2, double>(count, [&](int q) {
                const int i = index(q);
                double yo = GOY(i);
#pragma unroll
                for (int j = 0; j < DSB_RK_MAX_STAGES; ++j) if (j < ns) yo = GDF(j, i) * bf[j] + yo;
                return yo;
            }, store);
        } else {
            const double al1 = theta - 1.0, be1 = 1.0 - 2.0 * theta;
            const double al2 = 1.0 - theta, be2 = theta * (theta - 1.0);
            band_for<U2, double>(count, [&](int q) {
                const int i = index(q);
                const double u0 = GOY(i), u1 = GY(i);
                double v = u1;
                v -= u0;
                v = al1 * GDF(0, i) + be1 * v;
                v = theta * GDF(ns - 1, i) + v;
                v = al2 * u0 + be2 * v;
                v = theta * u1 + v;
                return v;
            }, store);
        }
    };
    auto interpolate_to = [&](double tq, auto&& store) { interpolate_range(tq, [](int q) { return q; }, N, store); };
    // into the (free) Newton residual vector, for the output and root functions: only the components they read when
    // the equations declare them (dsb_math.h: dsb_model_ndep)
    constexpr int NDEP = dsb_model_ndep<M>::value;
    auto interpolate_for_functions = [&](double tq) {
        if constexpr (NDEP > 0) interpolate_range(tq, [](int q) { return M::dep(q); }, NDEP, [&](int q, double yo) { GDL(M::dep(q)) = yo; });
        else interpolate_to(tq, [&](int i, double yo) { GDL(i) = yo; });
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
        const unsigned m_idle = __ballot_sync(0xffffffffu, state == R_IDLE);
        if (m_idle == 0xffffffffu) break;
        const int n_active = 32 - __popc(m_idle);
        const int n_slow = __popc(__ballot_sync(0xffffffffu, state == R_ERRTEST || state == R_JAC || state == R_ACCEPT));
        const bool run_slow = n_slow > 0 && (n_slow >= quorum || 2 * n_slow >= n_active);

        // ================= FINISH ===============================================================================
        if (__any_sync(0xffffffffu, state == R_FINISH) && state == R_FINISH) {
            bb.status[inst] = fin_status;
            bb.fin_t[inst] = t; bb.fin_h[inst] = h_state; bb.fin_order[inst] = pa.rk.order;
#pragma unroll
            for (int k = 0; k < DSB_NSTATS; ++k) bb.stats[(int64_t)k * B + inst] = st.v[k];
            if (NR > 0) { bb.ncols[inst] = col; bb.root_idx[inst] = root_found; }
            state = R_FETCH;
        }
        // ================= FETCH: next instance; new_without_initialise, set_step_size, Rk::_new + Sdirk::_new ==
        if (__any_sync(0xffffffffu, state == R_FETCH) && state == R_FETCH) {
            inst = (int64_t)atomicAdd(work_counter, 1ull);
            if (inst >= B) {
                state = R_IDLE;
            } else if (!M::HAS_MASS || bb.status[inst] == DSB_STATUS_OK) {     // else: consistent initialisation failed, keep its status
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = bb.params[(int64_t)j * B + inst];
#pragma unroll
                for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = 0;
                t = pa.t0; old_t = t;
                if constexpr (M::HAS_MASS) {
                    // singular mass: y, dy after set_consistent and the counters so far come from dsb_band_init_kernel
#pragma unroll
                    for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = bb.stats[(int64_t)k * B + inst];
                    band_for<U4, BandR2>(N, [&](int i) { return BandR2{bb.y0[(int64_t)i * B + inst], bb.dy0[(int64_t)i * B + inst]}; },
                                         [&](int i, const BandR2& r) { GY(i) = r.a; GDY(i) = r.b; });
                } else {
                    // y = init(p, t0); dy = f(y, t0)     (state.rs:1086-1124)
#pragma unroll UN
                    for (int i = 0; i < N; ++i) GY(i) = M::init_i(i, pl, pa.t0);
                    band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vY, pl, pa.t0); }, [&](int i, double r) { GDY(i) = r; });
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                }
                // set_step_size (state.rs:1209-1277), solver order = the tableau's
                {
                    const bool is_neg_h = pa.h0 < 0.0;
                    const double d0 = dsb_sqrt(weighted_norm(Lay::O_Y, Lay::O_Y));
                    const double d1 = dsb_sqrt(weighted_norm(Lay::O_DY, Lay::O_Y));
                    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * DSB_DIV(d0, d1);
                    band_for<U4, double>(N, [&](int i) { return is_neg_h ? (GDY(i) * (-h0) + GY(i)) : (GDY(i) * h0 + GY(i)); },
                                         [&](int i, double r) { GTMP(i) = r; });
                    const double t1 = is_neg_h ? pa.t0 - h0 : pa.t0 + h0;
                    band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vTMP, pl, t1) - GDY(i); }, [&](int i, double r) { GDL(i) = r; });
                    st.v[DSB_STAT_RHS_CALLS] += 1;
                    const double d2 = DSB_DIV(dsb_sqrt(weighted_norm(Lay::O_DL, Lay::O_Y)), dsb_abs(h0));
                    double max_d = d2;
                    if (max_d < d1) max_d = d1;
                    double h1;
                    if (max_d < 1e-15) { h1 = h0 * 1e-3; if (h1 < 1e-6) h1 = 1e-6; }
                    else h1 = dsb_pow(DSB_DIV(0.01, max_d), DSB_DIV(1.0, 1.0 + (double)pa.rk.order));
                    h_state = 100.0 * h0;
                    if (h_state > h1) h_state = h1;
                    if (is_neg_h) h_state = -h_state;
                }
                band_for<U4, double>(N, [&](int i) { return GY(i); }, [&](int i, double r) {
                    GOY(i) = r; GPHI(i) = 0.0;
#pragma unroll
                    for (int j = 0; j < DSB_RK_MAX_STAGES; ++j) GDF(j, i) = 0.0;
                });
                ju.init(1.0);
                ju.update_jacobian(h_state);
                ju.update_rhs_jacobian(h_state);
                conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
                op_h = h_state;
                jacobian_is_stale = true; is_jacobian_set = false;
                has_tstop = false; tstop = 0.0; has_prev_error = false; prev_error_norm = 0.0;
                first = true; reached = false; col = 0;
                if constexpr (NR > 0) {                         // Rk::_new: root_finder.init(root_fn, state.y, state.t)
                    M::root(vY, pl, t, rf.g0);
                    rf.t0 = t; root_found = -1;
                }
                state = R_TSTOP;
            }
        }

        // ================= ERRTEST: embedded error estimate, step-size factor, accept / reject ==================
        // (sdirk.rs:474-529, runge_kutta.rs:783-800, 466-495)
        if (run_slow && state == R_ERRTEST) {
            {
                double dco[DSB_RK_MAX_STAGES];
#pragma unroll
                for (int j = 0; j < DSB_RK_MAX_STAGES; ++j) dco[j] = (j < ns) ? pa.rk.d[j] : 0.0;
                band_for<U2, double>(N, [&](int k) {
                    double e = GDF(0, k) * dco[0];
#pragma unroll
                    for (int j = 1; j < DSB_RK_MAX_STAGES; ++j) if (j < ns) e = GDF(j, k) * dco[j] + e;
                    return e;
                }, [&](int k, double e) { if constexpr (M::HAS_MASS) GTMP(k) = e; else GDL(k) = e; });
            }
            if constexpr (M::HAS_MASS) {
                // error <- M error: column sweep, the first product is assigned (matrix gemv with beta = 0)
                band_for<U2, double>(N, [&](int k) {
                    double e = -0.0;                                  // (-0.0) + x == x
#pragma unroll
                    for (int d = -KL; d <= KU; ++d) {
                        const int j = k + d;
                        if (j >= 0 && j < N) e = GM(j, KU - d) * GTMP(j) + e;
                    }
                    return e;
                }, [&](int k, double e) { GDL(k) = e; });
            }
            if (!BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, Lay::O_DL)) {
                finish(DSB_STATUS_LU_SOLVE_FAILED);
            } else {
                const double e = weighted_norm(Lay::O_DL, Lay::O_Y);            // weights from state.y
                error_norm = (0.0 < e) ? e : 0.0;
                const double maxiter = (double)conv.max_iter;
                const double niter = (double)conv.niter;
                const double safety_factor = DSB_DIV(2.0 * maxiter + 1.0, 2.0 * maxiter + niter);
                const double safety = 0.9 * safety_factor;
                const double order_f = (double)(pa.rk.order + 1);
                const double ki = DSB_DIV(pa.opt.pi_control_integral, order_f);
                const bool p_only = pa.opt.pi_control_proportional == 0.0 || !has_prev_error;
                const double kp = p_only ? 0.0 : DSB_DIV(pa.opt.pi_control_proportional, order_f);
                double raw = dsb_pow(error_norm, p_only ? -ki : -(ki + kp));
                if (!p_only) raw = raw * dsb_pow(prev_error_norm, kp);
                double f = safety * raw;
                if (f > pa.opt.max_timestep_shrink && f < pa.opt.min_timestep_growth) f = 1.0;
                if (f < pa.opt.min_timestep_shrink) f = pa.opt.min_timestep_shrink;
                if (f > pa.opt.max_timestep_growth) f = pa.opt.max_timestep_growth;
                factor = f;
                if (error_norm < 1.0) {
                    const double new_h = h * factor;
                    if (factor != 1.0) conv.eta = pa.tab.eta_reset_timestep;
                    op_h = new_h;
                    jac_h = new_h; jac_kind = DSB_STEP_SUCCESS; after_jac = R_ACCEPT;
                    state = R_JAC;
                } else {
                    h *= factor;
                    conv.eta = pa.tab.eta_reset_timestep;
                    op_h = h;
                    jac_h = h; jac_kind = DSB_ERROR_TEST_FAIL; after_jac = R_ATTEMPT;
                    state = R_JAC;
                }
            }
        }

        // ================= JAC: Sdirk::jacobian_updates(h, kind) / the lazy first reset_jacobian ================
        if (run_slow && state == R_JAC) {
            bool do_factor = false;
            double t_jac = t;
            if (jac_kind == DSB_KIND_LAZY) {
                do_factor = true;
                t_jac = t_stage;
                st.record_linear_solver_setup(DSB_CHECKPOINT);
            } else if (ju.check_rhs_jacobian_update<DsbDivShared>(pa.opt, jac_h, jac_kind)) {
                jacobian_is_stale = true;
                ju.update_rhs_jacobian(jac_h);
                ju.update_jacobian(jac_h);
                do_factor = true;
            } else if (ju.check_jacobian_update<DsbDivShared>(pa.opt, jac_h, jac_kind)) {
                ju.update_jacobian(jac_h);
                do_factor = true;
            }
            if (do_factor) {
                if (jac_kind != DSB_KIND_LAZY) {
                    conv.eta = pa.tab.eta_reset;
                    st.record_linear_solver_setup(jac_kind);
                }
                if (jacobian_is_stale) {
                    // df/dy at phi + c * state.y with the phi left over from the last stage (op/sdirk.rs:265-276), one
                    // jac_mul per colour, scattered through the sparsity pattern into band storage
                    band_for<U4, double>(N, [&](int i) { return cg * GY(i) + GPHI(i); }, [&](int i, double r) { GTMP(i) = r; });
                    st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
                    for (int e = 0; e < LDJ * N; ++e) G(Lay::O_J + e) = 0.0;
                    const bool one_colour_per_column = pa.ncolors == N;
#pragma unroll 1
                    for (int cc = 0; cc < pa.ncolors; ++cc) {
                        const BandColourSeed seed{meta.colmeta, cc};
                        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
                        const int i0 = one_colour_per_column ? (cc - KU < 0 ? 0 : cc - KU) : 0;
                        const int i1 = one_colour_per_column ? (cc + KL > N - 1 ? N - 1 : cc + KL) : N - 1;
                        band_for<U2, double>(i1 - i0 + 1, [&](int q) { return M::jac_mul_i(i0 + q, vTMP, pl, t_jac, seed); },
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
#pragma unroll 1
                        for (int j = 0; j < N; ++j) {
                            const BandUnitVec ej{j};
#pragma unroll
                            for (int r = 0; r < LDJ; ++r) {
                                const int i = j + r - KU;
                                GM(j, r) = (i >= 0 && i < N) ? M::mass_i(i, ej, pl, t_jac, 0.0, 0.0) : 0.0;
                            }
                        }
                    }
                    jacobian_is_stale = false;
                }
                // A = M - (c h) J (op/sdirk.rs:277-292) in band storage with kl extra rows for the fill-in
                const double beta = -(cg * op_h);
                band_for<U2, BandRBand<LDAB>>(N, [&](int j) {
                    BandRBand<LDAB> a;
#pragma unroll
                    for (int r = 0; r < LDAB; ++r) {
                        const int i = j + r - KV;
                        double v = 0.0;
                        if (r >= KL && i >= 0 && i < N) {
                            if constexpr (M::HAS_MASS) v = GJ(j, r - KL) * beta + GM(j, r - KL);
                            else v = GJ(j, r - KL) * beta + ((i == j) ? 1.0 : 0.0);
                        }
                        a.v[r] = v;
                    }
                    return a;
                }, [&](int j, const BandRBand<LDAB>& a) {
#pragma unroll
                    for (int r = 0; r < LDAB; ++r) GAB(j, r) = a.v[r];
                });
                BLU::factor(g, LS, Lay::O_LU, Lay::O_PIV);
                is_jacobian_set = true;
            }
            state = after_jac;
            if (jac_kind != DSB_KIND_LAZY && jac_kind != DSB_STEP_SUCCESS) {
                // the failure paths continue after jacobian_updates (sdirk.rs:464-471, 524-529)
                has_prev_error = false;
                if (jac_kind == DSB_ERROR_TEST_FAIL) {
                    nattempts += 1;
                    st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
                    if (nattempts >= pa.opt.max_error_test_failures) finish(DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES);
                    else if (dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
                } else {
                    st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] += 1;
                    if (st.v[DSB_STAT_NONLINEAR_SOLVER_FAILS] > pa.opt.max_nonlinear_solver_failures)
                        finish(DSB_STATUS_TOO_MANY_NONLINEAR_FAILURES);
                    else if (dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
                }
            }
        }

        // ================= ACCEPT: rest of the accepted path + Rk::step_accepted (runge_kutta.rs:894-960) ========
        if (run_slow && state == R_ACCEPT) {
            ju.step();
            has_prev_error = true; prev_error_norm = error_norm;
            const double new_h = h * factor;
            const double inv_h = 1.0 / h;
            old_t = t;
            t = t + h;
            h_state = new_h;
            band_for<U2, BandRCols>(N, [&](int i) {
                BandRCols r;
                r.v[0] = GOY(i);                             // old_state.y held the last stage value
                r.v[1] = GY(i);                              // swap: old_state <- previous state
                r.v[2] = GX(i) * inv_h;                      // old_state.dy *= 1/h, then swapped in
                return r;
            }, [&](int i, const BandRCols& r) { GY(i) = r.v[0]; GOY(i) = r.v[1]; GDY(i) = r.v[2]; });
            st.v[DSB_STAT_STEPS] += 1;
            state = R_TSTOP;
        }

        // ================= TSTOP: set_stop_time (first) / handle_tstop after an accepted step ===================
        if (__any_sync(0xffffffffu, state == R_TSTOP) && state == R_TSTOP) {
            int r = 0;
            int next = first ? R_STEP : R_OUTPUT;
            bool stopped_on_root = false;
            bool reset_now = false;            // a reset was applied at a root: the stop time is set again, then R_STEP
            if constexpr (NR > 0) {
                // check for a root within the accepted step (runge_kutta.rs:935-948), before the stop time is handled;
                // the interpolated state of the secant iteration goes to the (free) Newton residual
                if (!first) {   // also in the step()/interpolate() loop of the reference's harness (free_running), which returns interpolate(t_root) and ends (ode_solver/mod.rs:134-141)
                    double t_root = t;
                    stopped_on_root = rf.check_root(t, [&](double (&gv)[NR]) { M::root(vY, pl, t, gv); },
                                                    [&](double t_mid, double (&gv)[NR]) {
                                                        interpolate_for_functions(t_mid);
                                                        M::root(vDL, pl, t_mid, gv);
                                                    }, t_root, root_found);
                    if (stopped_on_root) {
                        // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root)
                        // (runge_kutta.rs:396-434), then the state at the root in the next column (method.rs:493-503)
                        while (!free_running && col < nt && bb.t_eval[col] <= t_root) {
                            write_column(bb.t_eval[col], col);
                            ++col;
                        }
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (sdirk.rs:368-374 -> state.rs:279-306:
                                // y <- reset(y, t), dy <- f(y, t)), set_stop_time(final_time) and on with the
                                // integration -- or TstopReached.  Step size, Jacobian and LU stay; Rk::start_step
                                // (runge_kutta.rs:446-464) re-initialises the root finder and sets the stop time again.
                                interpolate_to(t_root, [&](int i, double yo) { GDL(i) = yo; });
                                t = t_root;
                                band_for<U2, double>(N, [&](int i) { return M::reset_i(i, vDL, pl, t); }, [&](int i, double v) { GY(i) = v; });
                                band_for<U2, double>(N, [&](int i) { return M::rhs_i(i, vY, pl, t); }, [&](int i, double v) { GDY(i) = v; });
                                st.v[DSB_STAT_RHS_CALLS] += 1;
                                root_found = -1;
                                ended = false;
                                if (t < bb.t_eval[nt - 1]) {
                                    has_tstop = true; tstop = bb.t_eval[nt - 1];
                                    r = handle_tstop(tstop);                       // method.rs:792
                                    if (r == 0) {
                                        M::root(vY, pl, t, rf.g0);                 // start_step: root_finder.init
                                        rf.t0 = t;
                                        r = handle_tstop(tstop);                   // start_step: set_stop_time(tstop)
                                    }
                                    if (r == 1) r = -DSB_STATUS_STOP_TIME_AT_CURRENT;
                                    stopped_on_root = false; reset_now = true;
                                    next = R_STEP;
                                } else finish(DSB_STATUS_OK);                      // TstopReached
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
            if (first) {
                if (free_running) next = R_OUTPUT;
                else {
                    has_tstop = true; tstop = bb.t_eval[nt - 1];
                    r = handle_tstop(tstop);
                    if (r == 1) r = -DSB_STATUS_STOP_TIME_AT_CURRENT;
                }
            } else if (has_tstop && !stopped_on_root && !reset_now) {
                r = handle_tstop(tstop);
                if (r == 1) { reached = true; has_tstop = false; }
            }
            if (stopped_on_root || state == R_FINISH) {
                // the lane is on its way to FINISH
            } else if (r < 0) finish(-r);
            else state = next;
            first = false;
        }

        // ================= OUTPUT: dense output (method.rs:761-764, 822-848; runge_kutta.rs:1080-1127) ==========
        if (__any_sync(0xffffffffu, state == R_OUTPUT) && state == R_OUTPUT) {
            int status = DSB_STATUS_OK;
            while (col < nt) {
                const double tq = bb.t_eval[col];
                if (free_running ? (dsb_abs(t) < dsb_abs(tq)) : !(tq <= t)) break;
                const bool is_forward = h_state > 0.0;
                if ((is_forward && (tq > t || tq < old_t)) || (!is_forward && (tq < t || tq > old_t))) {
                    status = DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT; break;
                }
                write_column(tq, col);
                ++col;
            }
            if (status != DSB_STATUS_OK) finish(status);
            else if (free_running ? (col >= nt) : reached) finish(DSB_STATUS_OK);
            else state = R_STEP;
        }

        // ================= STEP: start of Sdirk::step (sdirk.rs:415-431) ========================================
        if (__any_sync(0xffffffffu, state == R_STEP) && state == R_STEP) {
            h = h_state;
            if (dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
            else {
                op_h = h;
                nattempts = 0; updated_jacobian = false;
                state = R_ATTEMPT;
            }
        }
        // ================= ATTEMPT: start_step_attempt (runge_kutta.rs:505-535) =================================
        if (__any_sync(0xffffffffu, state == R_ATTEMPT) && state == R_ATTEMPT) {
            if (start == 1) {
                band_for<U4, double>(N, [&](int k) { return h * GDY(k); }, [&](int k, double r) { GDF(0, k) = r; });
            }
            stage = start;
            state = R_STAGE;
        }
        // ================= STAGE: set_phi + predict_stage_sdirk (runge_kutta.rs:645-665) ========================
        if (__any_sync(0xffffffffu, state == R_STAGE) && state == R_STAGE) {
            const int i = stage;
            t_stage = t + pa.rk.c[i] * h;
            double aco[DSB_RK_MAX_STAGES];
#pragma unroll
            for (int j = 0; j < DSB_RK_MAX_STAGES; ++j) aco[j] = (j < i) ? pa.rk.a[j * ns + i] : 0.0;
            double al = 0.0, be = 0.0;
            if (i >= 2) {
                const double cc = DSB_DIV(pa.rk.c[i] - pa.rk.c[i - 2], pa.rk.c[i - 1] - pa.rk.c[i - 2]);
                al = -cc; be = 1.0 + cc;
            }
            band_for<U2, BandR2>(N, [&](int k) {
                double ph = GY(k);
#pragma unroll
                for (int j = 0; j < DSB_RK_MAX_STAGES - 1; ++j) if (j < i) ph = GDF(j, k) * aco[j] + ph;
                double x;
                if (i == 0) x = h * GDY(k);
                else if (i == 1) x = GDF(0, k);
                else x = al * GDF(i - 2, k) + be * GDF(i - 1, k);
                return BandR2{ph, x};
            }, [&](int k, const BandR2& r) { GPHI(k) = r.a; GX(k) = r.b; });
            conv.reset();
            if (!is_jacobian_set) { jac_kind = DSB_KIND_LAZY; after_jac = R_NEWTON; state = R_JAC; }
            else state = R_NEWTON;
        }

        // ================= NEWTON: one iteration on F(x) = M x - h f(phi + c x) =================================
#pragma unroll 1
        for (int pass = 0; pass < newton_passes; ++pass) {
        if (!__any_sync(0xffffffffu, state == R_NEWTON)) break;
        if (state == R_NEWTON) {
            band_for<U4, double>(N, [&](int i) { return cg * GX(i) + GPHI(i); }, [&](int i, double r) { GTMP(i) = r; });
            const double beta = -op_h;
            band_for<U2, double>(N, [&](int i) {
                const double f = M::rhs_i(i, vTMP, pl, t_stage);
                if constexpr (M::HAS_MASS) return M::mass_i(i, vX, pl, t_stage, beta, f);     // gemv_inplace: y = M x + beta y
                else return GX(i) + beta * f;
            }, [&](int i, double r) { GDL(i) = r; });
            st.v[DSB_STAT_RHS_CALLS] += 1;
            if (!BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, Lay::O_DL)) {
                newton_ok = false; state = R_POST;
            } else {
                double acc = 0.0;
                band_for<U4, BandR2>(N, [&](int i) {
                    const double dl = GDL(i);
                    return BandR2{GX(i) - dl, DSB_DIV(dl, dsb_abs(GY(i)) * pa.rtol + meta.atol[i])};    // weights from state.y
                }, [&](int i, const BandR2& r) { GX(i) = r.a; acc += r.b * r.b; });
                const double norm = dsb_sqrt(DSB_DIV(acc, (double)N));
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
                if (conv.niter == 1) { conv.has_old_norm = true; conv.old_norm = norm; }
                if (s == LANE_CONVERGED) { newton_ok = true; state = R_POST; }
                else if (s == LANE_DIVERGED || conv.niter >= conv.max_iter) { newton_ok = false; state = R_POST; }
            }
        }
        }

        // ================= POST: a stage's Newton solve ended (runge_kutta.rs:674-679, sdirk.rs:436-472) ========
        if (__any_sync(0xffffffffu, state == R_POST) && state == R_POST) {
            st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
            if (newton_ok) {
                const int i = stage;
                band_for<U4, BandR2>(N, [&](int k) {
                    const double x = GX(k);
                    return BandR2{cg * x + GPHI(k), x};      // get_f_eval: stage value
                }, [&](int k, const BandR2& r) { GOY(k) = r.a; GDF(i, k) = r.b; });
                stage = i + 1;
                state = (stage < ns) ? R_STAGE : R_ERRTEST;
            } else {
                if (!updated_jacobian) {
                    updated_jacobian = true;
                    jac_kind = DSB_FIRST_CONVERGENCE_FAIL;
                } else {
                    h *= 0.3;
                    conv.eta = pa.tab.eta_reset_timestep;
                    op_h = h;
                    jac_kind = DSB_SECOND_CONVERGENCE_FAIL;
                }
                jac_h = h; after_jac = R_ATTEMPT;
                state = R_JAC;
            }
        }
    }
#undef SMW
#undef G
#undef GDF
#undef GY
#undef GDY
#undef GOY
#undef GPHI
#undef GX
#undef GDL
#undef GTMP
#undef GJ
#undef GAB
#undef GM
#undef DSB_DIV
}
