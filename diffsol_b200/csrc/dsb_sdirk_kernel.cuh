// dsb_sdirk_kernel.cuh -- `problem.tr_bdf2::<LS>()?.solve_dense(t_eval)` / `esdirk34` for every instance
// of a batch: the same execution model as dsb_bdf_kernel.cuh (one lane per instance, per-lane state
// machine, warp-level block scheduler, persistent work-fetching grid), applied to the (E)SDIRK loop.
//
//   FETCH -> TSTOP(first) -> STEP -> ATTEMPT -> STAGE(i) -> [JAC lazily] -> NEWTON* -> POST
//         -> STAGE(i+1) ... -> ERRTEST -> (accept) JAC(StepSuccess) -> ACCEPT -> TSTOP -> OUTPUT -> STEP ...
//                                      -> (reject) JAC(ErrorTestFail) -> ATTEMPT ...
//         POST (Newton failed) -> JAC(First/SecondConvergenceFail) -> ATTEMPT ...
//
// Restated functions (paths relative to /root/reference/crates/diffsol/src):
//   Sdirk::_new, jacobian_updates, step     ode_solver/sdirk.rs:178-304, 409-543
//   Rk::_new, start_step_attempt, do_stage_sdirk, predict_stage_sdirk, error_norm, factor, solve_fail,
//   error_test_fail, step_accepted, handle_tstop, set_stop_time, interpolate_inplace
//                                            ode_solver/runge_kutta.rs:100-175, 431-441, 466-535, 610-981, 1080-1127
//   pi_controller_raw                        ode_solver/runge_kutta.rs:1313-1335
//   SdirkCallable (set_phi, call_inplace, jacobian_inplace, get_f_eval)   op/sdirk.rs:157-292
//   newton_iteration + NoLineSearch, Convergence   crates/diffsol-nl/src/newton.rs:13-36, line_search.rs:48-69,
//                                            convergence.rs:64-139
// Reference quirks kept on purpose: df/dy is evaluated at phi + c * state.y with the phi left over from
// the last stage (op/sdirk.rs:265-276); JacobianUpdate is fed h, not a_d * h; the first LU is set up lazily
// inside the first stage and counted as a Checkpoint; the accept test is strict; the safety factor uses the
// Newton iteration count of the last stage only; the error estimate is filtered through the LU.
#pragma once
#include "dsb_lane.cuh"
#include "dsb_init_kernel.cuh"      // lane_consistent_solve: consistent sensitivities of a DAE
#include "dsb_roots.cuh"

enum dsb_rk_lane_state {
    R_FETCH = 0, R_FINISH, R_ERRTEST, R_JAC, R_ACCEPT, R_TSTOP, R_OUTPUT, R_STEP, R_ATTEMPT, R_STAGE, R_NEWTON, R_POST, R_IDLE
};
#define DSB_KIND_LAZY 5          // the first reset_jacobian, inside do_stage_sdirk (runge_kutta.rs:661-665)
#define DSB_RK_MAX_STAGES 4

template <class M>
struct SdirkLayout {
    static constexpr int N = M::N, NP = M::NP;
    static constexpr int O_DIFF = 0;                                // diff[DSB_RK_MAX_STAGES][N]: x_i = h k_i
    static constexpr int O_J = O_DIFF + DSB_RK_MAX_STAGES * N;      // rhs_jac[col][row]
    static constexpr int O_M = O_J + N * N;                         // mass_jac[col][row] (DAE only)
    static constexpr int O_LU = O_M + (M::HAS_MASS ? N * N : 0);    // LU factors
    static constexpr int O_Y = O_LU + N * N;                        // state.y
    static constexpr int O_DY = O_Y + N;                            // state.dy
    static constexpr int O_OY = O_DY + N;                           // old_state.y (last stage value / previous step)
    static constexpr int O_PHI = O_OY + N;                          // SdirkCallable.phi
    static constexpr int O_P = O_PHI + N;                           // parameters
    static constexpr int O_ST = O_P + (NP > 0 ? NP : 1);            // statistics, two int32 per word
    // forward sensitivities (DsbWithSens<M> only; Rk: runge_kutta.rs:46, 518-523, 691-745, 812-822, 917-920): per parameter the
    // stage increments, state.s / state.ds, old_state.s (the stage value, then the previous step's s) and the column of f_p at
    // the stage value; phi of the sensitivity residual
    static constexpr bool SENS = dsb_model_sens_on<M>::value;
    static constexpr int O_SSD = O_ST + (DSB_NSTATS + 1) / 2;       // sdiff[NP][DSB_RK_MAX_STAGES][N]
    static constexpr int O_SS = O_SSD + NP * DSB_RK_MAX_STAGES * N; // state.s[NP][N]
    static constexpr int O_SDS = O_SS + NP * N;                     // state.ds[NP][N]
    static constexpr int O_SOS = O_SDS + NP * N;                    // old_state.s[NP][N]
    static constexpr int O_SFP = O_SOS + NP * N;                    // f_p e_q at the stage value [NP][N]
    static constexpr int O_SPH = O_SFP + NP * N;                    // phi of SdirkCallable<SensEquations>
    static constexpr int WORDS = SENS ? O_SPH + N : O_SSD;
    static constexpr int THREADS = LaneBlockShape<WORDS, N>::THREADS;
    static constexpr int MAXNREG = LaneBlockShape<WORDS, N>::MAXNREG;
};

template <class M>
__global__ void __maxnreg__(SdirkLayout<M>::MAXNREG)
dsb_sdirk_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa, const __grid_constant__ DsbBatchBuffers bb,
                             unsigned long long* __restrict__ work_counter) {
    constexpr int N = M::N;
    constexpr int NP = M::NP;
    static_assert(N <= 16, "pivots are packed 4 bits per row");
    typedef SdirkLayout<M> Lay;
    extern __shared__ double dsb_lane_smem[];
    double* const sm = dsb_lane_smem + threadIdx.x;
#define SM(w) sm[(w) * Lay::THREADS]
#define DSB_DIV(a, b) DsbDivInline::div((a), (b))      // this kernel fits the instruction cache (dsb_math.h)
#define SDF(j, i) SM(Lay::O_DIFF + (j) * N + (i))
#define SJ(j, i) SM(Lay::O_J + (j) * N + (i))
#define SMM(j, i) SM(Lay::O_M + (j) * N + (i))
#define SLU(j, i) SM(Lay::O_LU + (j) * N + (i))
#define SY(i) SM(Lay::O_Y + (i))
#define SDY(i) SM(Lay::O_DY + (i))
#define SOY(i) SM(Lay::O_OY + (i))
#define SPHI(i) SM(Lay::O_PHI + (i))
#define SP(i) SM(Lay::O_P + (i))
#define SSD(q, j, i) SM(Lay::O_SSD + ((q) * DSB_RK_MAX_STAGES + (j)) * N + (i))
#define SSS(q, i) SM(Lay::O_SS + (q) * N + (i))
#define SDS(q, i) SM(Lay::O_SDS + (q) * N + (i))
#define SOS(q, i) SM(Lay::O_SOS + (q) * N + (i))
#define SFP(q, i) SM(Lay::O_SFP + (q) * N + (i))
#define SPHS(i) SM(Lay::O_SPH + (i))
    constexpr bool SENS = Lay::SENS;
    static_assert(!SENS || (dsb_model_has_sens<M>::value && dsb_model_nroots<M>::value == 0 &&
                            !dsb_model_nout<M>::has_out && !dsb_model_has_reset<M>::value),
                  "sensitivities: equations with sens_mul / init_sens, no root / output / reset functions");
    int eq = 0;            // sensitivities: which equation the Newton block is solving (0 = the state, q + 1 = sensitivity q)

    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const int quorum = pa.quorum;
    const int newton_passes = pa.newton_passes < 1 ? 1 : pa.newton_passes;
    const double eps = 2.220446049250313e-16;
    const int ns = pa.rk.s;
    const int start = (pa.rk.a[0] == 0.0) ? 1 : 0;        // skip_first_stage (runge_kutta.rs:286-288)
    const double cg = pa.rk.a[1 * ns + 1];                // Sdirk::gamma() = a(1, 1)

    // ---- per-lane registers ---------------------------------------------------------------------------
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
    st.v.base = reinterpret_cast<int*>(&SM(Lay::O_ST));
    unsigned long long piv_packed = 0;
    double x_cur[N], wt[N];                                // Newton iterate (old_state.dy), norm weights from state.y
#pragma unroll
    for (int i = 0; i < N; ++i) { x_cur[i] = 0.0; wt[i] = 1.0; }
    int stage = 0, nattempts = 0, col = 0;
    bool updated_jacobian = false, newton_ok = false, first = true, reached = false;
    double t_stage = 0.0, factor = 1.0, error_norm = 0.0;
    int after_jac = R_NEWTON, jac_kind = DSB_CHECKPOINT;
    double jac_h = 0.0;
    int fin_status = DSB_STATUS_OK;
    auto finish = [&](int status) { fin_status = status; state = R_FINISH; };
    // root finding (nonlinear_solver/root.rs; runge_kutta.rs:43, 142-147, 935-948): only compiled for equations with roots
    constexpr int NR = dsb_model_nroots<M>::value;
    LaneRootFinder<(NR > 0 ? NR : 1), DsbDivInline> rf;
    rf.t0 = 0.0;
    int root_found = -1;
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
    // one column of the solve_dense result (dense_write_out, method.rs:822-848): the state, or -- for equations with an
    // output function (OdeEquations::out) -- out(y(tq), tq)
    // In the solve(final_time) form (DsbRagged<M>; write_out, method.rs:965-1000) the column goes, with its time, to the
    // instance's own run of the ragged result (writing pass only).
    constexpr bool RAG = dsb_model_ragged_on<M>::value;
    static_assert(!RAG || !dsb_model_has_reset<M>::value, "solve(final_time) form: no reset functions");
    auto write_column = [&](int column, double tq, const double (&yo)[N]) {
        if constexpr (dsb_model_nout<M>::has_out) {
            constexpr int NOUT = dsb_model_nout<M>::value;
            double pl_[NP > 0 ? NP : 1], o[NOUT];
#pragma unroll
            for (int j = 0; j < NP; ++j) pl_[j] = SP(j);
            M::out(yo, pl_, tq, o);
            if constexpr (RAG) {
                if (pa.ragged == 2) {
                    const int64_t at = bb.rag_off[inst] + column;
                    bb.rag_ts[at] = tq;
#pragma unroll
                    for (int k = 0; k < NOUT; ++k) bb.rag_ys[at * NOUT + k] = o[k];
                }
            } else {
#pragma unroll
            for (int k = 0; k < NOUT; ++k) bb.ys[((int64_t)column * NOUT + k) * B + inst] = o[k];
            }
        } else {
            (void)tq;
            if constexpr (RAG) {
                if (pa.ragged == 2) {
                    const int64_t at = bb.rag_off[inst] + column;
                    bb.rag_ts[at] = tq;
#pragma unroll
                    for (int i = 0; i < N; ++i) bb.rag_ys[at * N + i] = yo[i];
                }
            } else {
#pragma unroll
            for (int i = 0; i < N; ++i) bb.ys[((int64_t)column * N + i) * B + inst] = yo[i];
            }
        }
    };
    // interpolate_inplace (runge_kutta.rs:1080-1127; :962-981 beta dense output, :1004-1024 Hermite) on [old_t, t]
    auto interpolate = [&](double tq, double (&yo)[N]) {
        const double dt = t - old_t;
        const double theta = (dt == 0.0) ? 1.0 : DSB_DIV(tq - old_t, dt);
        if (pa.rk.has_beta) {
            const double th2 = theta * theta;
#pragma unroll
            for (int i = 0; i < N; ++i) yo[i] = SOY(i);
#pragma unroll 1
            for (int j = 0; j < ns; ++j) {
                double bf = pa.rk.beta[j] * theta;
                bf = pa.rk.beta[ns + j] * th2 + bf;
#pragma unroll
                for (int i = 0; i < N; ++i) yo[i] = SDF(j, i) * bf + yo[i];
            }
        } else {
            const double al1 = theta - 1.0, be1 = 1.0 - 2.0 * theta;
            const double al2 = 1.0 - theta, be2 = theta * (theta - 1.0);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double u0 = SOY(i), u1 = SY(i);
                double v = u1;
                v -= u0;
                v = al1 * SDF(0, i) + be1 * v;
                v = theta * SDF(ns - 1, i) + v;
                v = al2 * u0 + be2 * v;
                v = theta * u1 + v;
                yo[i] = v;
            }
        }
    };

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
    // ---- sensitivities (only instantiated for DsbWithSens<M>) ----
    // f_p e_q at (x, tq) for every parameter: SensRhs::update_state (sens_equations.rs:129-134)
    auto sens_update_state = [&](const double (&x)[N], double tq) {
        if constexpr (SENS) {
            double pl_[NP > 0 ? NP : 1], e[NP > 0 ? NP : 1], colv[N];
#pragma unroll
            for (int j = 0; j < NP; ++j) { pl_[j] = SP(j); e[j] = 0.0; }
#pragma unroll 1
            for (int q = 0; q < NP; ++q) {
#pragma unroll
                for (int j = 0; j < NP; ++j) e[j] = (j == q) ? 1.0 : 0.0;
                M::sens_mul(x, pl_, tq, e, colv);
#pragma unroll
                for (int i = 0; i < N; ++i) SFP(q, i) = colv[i];
            }
        }
    };
    // the start of stage `stage`'s Newton solve for sensitivity q (runge_kutta.rs:698-716): set_phi on sdiff[q] / state.s[q],
    // predict_stage_sdirk on state.ds[q] / sdiff[q]
    auto sens_stage_setup = [&](int q) {
        const int i = stage;
        double ph[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ph[k] = SSS(q, k);
#pragma unroll 1
        for (int j = 0; j < i; ++j) {
            const double aij = pa.rk.a[j * ns + i];
#pragma unroll
            for (int k = 0; k < N; ++k) ph[k] = SSD(q, j, k) * aij + ph[k];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) SPHS(k) = ph[k];
        if (i == 0) {
#pragma unroll
            for (int k = 0; k < N; ++k) x_cur[k] = h * SDS(q, k);
        } else if (i == 1) {
#pragma unroll
            for (int k = 0; k < N; ++k) x_cur[k] = SSD(q, 0, k);
        } else {
            const double cc = DSB_DIV(pa.rk.c[i] - pa.rk.c[i - 2], pa.rk.c[i - 1] - pa.rk.c[i - 2]);
            const double al = -cc, be = 1.0 + cc;
#pragma unroll
            for (int k = 0; k < N; ++k) x_cur[k] = al * SSD(q, i - 2, k) + be * SSD(q, i - 1, k);
        }
        conv.reset();
    };

    while (true) {
        // ---- warp-level block scheduler (see dsb_bdf_kernel.cuh) ------------------------------------------
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
            if (NR > 0 || RAG) bb.ncols[inst] = col;
            if (NR > 0) bb.root_idx[inst] = root_found;
            state = R_FETCH;
        }
        // ================= FETCH: next instance; Rk::_new + Sdirk::_new =========================================
        if (__any_sync(0xffffffffu, state == R_FETCH) && state == R_FETCH) {
            inst = (int64_t)atomicAdd(work_counter, 1ull);
            if (inst >= B) {
                state = R_IDLE;
            } else if (bb.status[inst] == DSB_STATUS_OK) {
#pragma unroll
                for (int j = 0; j < NP; ++j) SP(j) = bb.params[(int64_t)j * B + inst];
#pragma unroll
                for (int k = 0; k < DSB_NSTATS; ++k) st.v[k] = bb.stats[(int64_t)k * B + inst];
                t = pa.t0; h_state = bb.h0[inst]; old_t = t;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double yi = bb.y0[(int64_t)i * B + inst];
                    SY(i) = yi; SOY(i) = yi; SDY(i) = bb.dy0[(int64_t)i * B + inst]; SPHI(i) = 0.0;
                    wt[i] = dsb_abs(yi) * pa.rtol + pa.atol[i];
                }
#pragma unroll
                for (int j = 0; j < DSB_RK_MAX_STAGES; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) SDF(j, i) = 0.0;
                ju.init(1.0);
                ju.update_jacobian(h_state);
                ju.update_rhs_jacobian(h_state);
                conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
                op_h = h_state;
                jacobian_is_stale = true; is_jacobian_set = false;
                has_tstop = false; tstop = 0.0; has_prev_error = false; prev_error_norm = 0.0;
                first = true; reached = false; col = 0;
                if constexpr (NR > 0) {                         // Rk::_new: root_finder.init(root_fn, state.y, state.t)
                    double y0l[N], pl0[NP > 0 ? NP : 1];
#pragma unroll
                    for (int i = 0; i < N; ++i) y0l[i] = SY(i);
#pragma unroll
                    for (int j = 0; j < NP; ++j) pl0[j] = SP(j);
                    M::root(y0l, pl0, t, rf.g0);
                    rf.t0 = t; root_found = -1;
                }
                state = R_TSTOP;
                if constexpr (SENS) {
                    // RkState::new_with_sensitivities_and_consistent (state.rs:1032-1080) as in the Bdf kernel: s_q = (d y0 / d p)
                    // e_q, ds_q = J(y0) s_q + f_p e_q, for a DAE the InitOp solve on SensRhs; then Sdirk::new_augmented's
                    // jacobian_updates(h, Checkpoint) (sdirk.rs:252): the first LU is NOT the lazy one of the first stage
                    double y0l[N], pl0[NP > 0 ? NP : 1], e[NP > 0 ? NP : 1], sq[N], dsq[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) y0l[i] = SY(i);
#pragma unroll
                    for (int j = 0; j < NP; ++j) { pl0[j] = SP(j); e[j] = 0.0; }
                    sens_update_state(y0l, t);
#pragma unroll 1
                    for (int q = 0; q < NP; ++q) {
#pragma unroll
                        for (int j = 0; j < NP; ++j) e[j] = (j == q) ? 1.0 : 0.0;
                        M::init_sens(pl0, pa.t0, e, sq);
                        M::jac_mul(y0l, pl0, t, sq, dsq);
                        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll
                        for (int i = 0; i < N; ++i) { dsq[i] += SFP(q, i); SSS(q, i) = sq[i]; SDS(q, i) = dsq[i]; }
#pragma unroll 1
                        for (int j = 0; j < DSB_RK_MAX_STAGES; ++j)
#pragma unroll
                            for (int i = 0; i < N; ++i) SSD(q, j, i) = 0.0;
                    }
                    int ic_status = DSB_STATUS_OK;
                    if constexpr (M::HAS_MASS) {
                        LaneConvergence ic_conv;
                        ic_conv.tol = pa.opt.nonlinear_solver_tolerance;
                        ic_conv.eta = pa.tab.eta_reset;
                        ic_conv.max_iter = pa.opt.ic_max_newton_iterations;
                        ic_conv.reset();
#pragma unroll 1
                        for (int q = 0; q < NP && ic_status == DSB_STATUS_OK; ++q) {
#pragma unroll
                            for (int i = 0; i < N; ++i) { sq[i] = SSS(q, i); dsq[i] = SDS(q, i); }
                            ic_status = lane_consistent_solve<M>(pa, pl0, sq, dsq,
                                [&](const double (&x)[N], double (&out)[N]) {
                                    M::jac_mul(y0l, pl0, t, x, out);
                                    st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll
                                    for (int i = 0; i < N; ++i) out[i] += SFP(q, i);
                                },
                                [&](double (&J)[N][N]) {
                                    lane_jacobian_to<M>(pa, y0l, pl0, t, st, [&](int j, int i, double val) { J[j][i] = val; });
                                }, ic_conv, false, pa.opt.ic_use_linesearch != 0);
#pragma unroll
                            for (int i = 0; i < N; ++i) { SSS(q, i) = sq[i]; SDS(q, i) = dsq[i]; }
                        }
                    }
#pragma unroll 1
                    for (int q = 0; q < NP; ++q)
#pragma unroll
                        for (int i = 0; i < N; ++i) SOS(q, i) = SSS(q, i);          // old_state = state.clone()
                    eq = 0;
                    if (ic_status != DSB_STATUS_OK) finish(ic_status);
                    else { jac_kind = DSB_CHECKPOINT; jac_h = h_state; after_jac = R_TSTOP; state = R_JAC; }
                }
            }
        }

        // ================= ERRTEST: embedded error estimate, step-size factor, accept / reject ==================
        // (sdirk.rs:474-529, runge_kutta.rs:783-800, 466-495)
        if (run_slow && state == R_ERRTEST) {
            double err[N];
#pragma unroll
            for (int k = 0; k < N; ++k) err[k] = SDF(0, k) * pa.rk.d[0];
#pragma unroll 1
            for (int j = 1; j < ns; ++j) {
                const double dj = pa.rk.d[j];
#pragma unroll
                for (int k = 0; k < N; ++k) err[k] = SDF(j, k) * dj + err[k];
            }
            if (M::HAS_MASS) {
                double nx[N];
#pragma unroll
                for (int k = 0; k < N; ++k) nx[k] = err[k];
#pragma unroll
                for (int k = 0; k < N; ++k) err[k] = SMM(0, k) * nx[0];
#pragma unroll
                for (int j = 1; j < N; ++j)
#pragma unroll
                    for (int k = 0; k < N; ++k) err[k] = SMM(j, k) * nx[j] + err[k];
            }
            LaneLU<N> lu;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                lu.piv[j] = (int)((piv_packed >> (4 * j)) & 15ull);
#pragma unroll
                for (int i = 0; i < N; ++i) lu.a[j][i] = SLU(j, i);
            }
            if (!lu.solve(err)) {
                finish(DSB_STATUS_LU_SOLVE_FAILED);
            } else {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double term = DSB_DIV(err[i], wt[i]);                 // weights from state.y
                    acc += term * term;
                }
                const double e = DSB_DIV(acc, (double)N);
                error_norm = (0.0 < e) ? e : 0.0;
                if constexpr (SENS) {               // sdiff[q] . d, NOT filtered through the LU (runge_kutta.rs:812-822)
                    if (pa.sens_error_control) {
#pragma unroll 1
                        for (int q = 0; q < NP; ++q) {
                            double es[N];
#pragma unroll
                            for (int k = 0; k < N; ++k) es[k] = SSD(q, 0, k) * pa.rk.d[0];
#pragma unroll 1
                            for (int j = 1; j < ns; ++j) {
                                const double dj = pa.rk.d[j];
#pragma unroll
                                for (int k = 0; k < N; ++k) es[k] = SSD(q, j, k) * dj + es[k];
                            }
                            double accs = 0.0;
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double term = DSB_DIV(es[i], dsb_abs(SSS(q, i)) * pa.sens_rtol + pa.sens_atol[i]);
                                accs += term * term;
                            }
                            const double en = DSB_DIV(accs, (double)N);
                            error_norm = (error_norm < en) ? en : error_norm;
                        }
                    }
                }
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
            } else if (ju.check_rhs_jacobian_update(pa.opt, jac_h, jac_kind)) {
                jacobian_is_stale = true;
                ju.update_rhs_jacobian(jac_h);
                ju.update_jacobian(jac_h);
                do_factor = true;
            } else if (ju.check_jacobian_update(pa.opt, jac_h, jac_kind)) {
                ju.update_jacobian(jac_h);
                do_factor = true;
            }
            if (do_factor) {
                if (jac_kind != DSB_KIND_LAZY) {
                    conv.eta = pa.tab.eta_reset;
                    st.record_linear_solver_setup(jac_kind);
                }
                LaneLU<N> lu;
                double pl[NP > 0 ? NP : 1];
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = SP(j);
                if (jacobian_is_stale) {
                    double tmpv[N];                                      // set_tmp: phi + c * x with x = state.y
#pragma unroll
                    for (int i = 0; i < N; ++i) tmpv[i] = cg * SY(i) + SPHI(i);
                    lane_jacobian_to<M>(pa, tmpv, pl, t_jac, st, [&](int j, int i, double val) { SJ(j, i) = val; });
                    if (M::HAS_MASS) {
                        lane_mass_matrix<M>(pl, t_jac, lu.a);
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int i = 0; i < N; ++i) SMM(j, i) = lu.a[j][i];
                    }
                    jacobian_is_stale = false;
                }
                const double beta = -(cg * op_h);
#pragma unroll
                for (int j = 0; j < N; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double m_ji = M::HAS_MASS ? SMM(j, i) : ((i == j) ? 1.0 : 0.0);
                        lu.a[j][i] = SJ(j, i) * beta + m_ji;
                    }
                lu.factor();
                piv_packed = 0;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    piv_packed |= (unsigned long long)lu.piv[j] << (4 * j);
#pragma unroll
                    for (int i = 0; i < N; ++i) SLU(j, i) = lu.a[j][i];
                }
                is_jacobian_set = true;
            }
            state = after_jac;
            if (jac_kind != DSB_KIND_LAZY && jac_kind != DSB_STEP_SUCCESS && jac_kind != DSB_CHECKPOINT) {
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
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double y_new = SOY(i);                 // old_state.y held the last stage value
                SOY(i) = SY(i);                              // swap: old_state <- previous state
                SY(i) = y_new;
                // old_state.dy *= 1/h, then swapped in (the last stage's increment; with sensitivities x_cur has moved on)
                SDY(i) = (SENS ? SDF(ns - 1, i) : x_cur[i]) * inv_h;
                wt[i] = dsb_abs(y_new) * pa.rtol + pa.atol[i];
            }
            if constexpr (SENS) {
#pragma unroll 1
                for (int q = 0; q < NP; ++q) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double s_new = SOS(q, i);          // old_state.s held the last stage value
                        SOS(q, i) = SSS(q, i);
                        SSS(q, i) = s_new;
                        SDS(q, i) = SSD(q, ns - 1, i) * inv_h;
                    }
                }
            }
            st.v[DSB_STAT_STEPS] += 1;
            state = R_TSTOP;
        }

        // ================= TSTOP: set_stop_time (first) / handle_tstop after an accepted step ===================
        if (__any_sync(0xffffffffu, state == R_TSTOP) && state == R_TSTOP) {
            if constexpr (RAG) {                    // solve(final_time): the initial column, before the stop time is set (method.rs:900-901)
                if (first) {
                    double y0c[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) y0c[i] = SY(i);
                    write_column(0, t, y0c);
                    col = 1;
                }
            }
            int r = 0;
            int next = first ? R_STEP : R_OUTPUT;
            bool stopped_on_root = false;
            bool reset_now = false;            // a reset was applied at a root: the stop time is set again, then R_STEP
            if constexpr (NR > 0) {
                // check for a root within the accepted step (runge_kutta.rs:935-948), before the stop time is handled
                if (!first) {   // also in the step()/interpolate() loop of the reference's harness (free_running), which returns interpolate(t_root) and ends (ode_solver/mod.rs:134-141)
                    double pl[NP > 0 ? NP : 1], ys[N];
#pragma unroll
                    for (int j = 0; j < NP; ++j) pl[j] = SP(j);
#pragma unroll
                    for (int i = 0; i < N; ++i) ys[i] = SY(i);
                    double t_root = t;
                    stopped_on_root = rf.check_root(t, [&](double (&g)[NR]) { M::root(ys, pl, t, g); },
                                                    [&](double t_mid, double (&g)[NR]) {
                                                        double ymid[N];
                                                        interpolate(t_mid, ymid);
                                                        M::root(ymid, pl, t_mid, g);
                                                    }, t_root, root_found);
                    if (stopped_on_root) {
                        // fn solve_dense, RootFound (method.rs:774-805): the points up to the root, state_mut_back(t_root)
                        // (runge_kutta.rs:396-434), then the state at the root in the next column (method.rs:493-503)
                        double yo[N];
                        while (!RAG && !free_running && col < nt && bb.t_eval[col] <= t_root) {
                            interpolate(bb.t_eval[col], yo);
                            write_column(col, bb.t_eval[col], yo);
                            ++col;
                        }
                        interpolate(t_root, yo);
                        if (!free_running) t = t_root;      // state_mut_back; the harness loop leaves the state at the end of the step
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (sdirk.rs:368-374 -> state.rs:279-306:
                                // y <- reset(y, t), dy <- f(y, t)), then set_stop_time(final_time) and on with the
                                // integration -- or TstopReached.  The step size, the Jacobian and the LU stay as they
                                // are; Rk::start_step (runge_kutta.rs:446-464) finds the state mutated, re-initialises
                                // the root finder and sets the stop time once more.
                                double yr[N], dyr[N];
                                M::reset(yo, pl, t, yr);
                                M::rhs(yr, pl, t, dyr);
                                st.v[DSB_STAT_RHS_CALLS] += 1;
#pragma unroll
                                for (int i = 0; i < N; ++i) {
                                    SY(i) = yr[i]; SDY(i) = dyr[i];
                                    wt[i] = dsb_abs(yr[i]) * pa.rtol + pa.atol[i];
                                }
                                root_found = -1;
                                ended = false;
                                if (t < bb.t_eval[nt - 1]) {
                                    has_tstop = true; tstop = bb.t_eval[nt - 1];
                                    r = handle_tstop(tstop);                       // method.rs:792
                                    if (r == 0) {
                                        M::root(yr, pl, t, rf.g0);                 // start_step: root_finder.init
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
                            if (RAG || col < nt) {
                                write_column(col, t_root, yo);
                                ++col;
                            }
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
            if constexpr (RAG) {                    // solve(final_time): (state.t, state.y) after every step (method.rs:907-921)
                double yc[N];
#pragma unroll
                for (int i = 0; i < N; ++i) yc[i] = SY(i);
                write_column(col, t, yc);
                ++col;
            }
            while (!RAG && col < nt) {
                const double tq = bb.t_eval[col];
                if (free_running ? (dsb_abs(t) < dsb_abs(tq)) : !(tq <= t)) break;
                const bool is_forward = h_state > 0.0;
                if ((is_forward && (tq > t || tq < old_t)) || (!is_forward && (tq < t || tq > old_t))) {
                    status = DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT; break;
                }
                double yo[N];
                interpolate(tq, yo);
                write_column(col, tq, yo);
                if constexpr (SENS) {               // interpolate_sens (runge_kutta.rs:1237-1310) on (old_state.s, state.s, sdiff)
                    const double dt = t - old_t;
                    const double theta = (dt == 0.0) ? 1.0 : DSB_DIV(tq - old_t, dt);
#pragma unroll 1
                    for (int q = 0; q < NP; ++q) {
                        if (pa.rk.has_beta) {
                            const double th2 = theta * theta;
#pragma unroll
                            for (int i = 0; i < N; ++i) yo[i] = SOS(q, i);
#pragma unroll 1
                            for (int j = 0; j < ns; ++j) {
                                double bf = pa.rk.beta[j] * theta;
                                bf = pa.rk.beta[ns + j] * th2 + bf;
#pragma unroll
                                for (int i = 0; i < N; ++i) yo[i] = SSD(q, j, i) * bf + yo[i];
                            }
                        } else {
                            const double al1 = theta - 1.0, be1 = 1.0 - 2.0 * theta;
                            const double al2 = 1.0 - theta, be2 = theta * (theta - 1.0);
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double u0 = SOS(q, i), u1 = SSS(q, i);
                                double v = u1;
                                v -= u0;
                                v = al1 * SSD(q, 0, i) + be1 * v;
                                v = theta * SSD(q, ns - 1, i) + v;
                                v = al2 * u0 + be2 * v;
                                v = theta * u1 + v;
                                yo[i] = v;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < N; ++i) bb.ss[(((int64_t)col * NP + q) * N + i) * B + inst] = yo[i];
                    }
                }
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
#pragma unroll
                for (int k = 0; k < N; ++k) SDF(0, k) = h * SDY(k);
                if constexpr (SENS) {
#pragma unroll 1
                    for (int q = 0; q < NP; ++q)
#pragma unroll
                        for (int k = 0; k < N; ++k) SSD(q, 0, k) = h * SDS(q, k);
                }
            }
            stage = start;
            eq = 0;
            state = R_STAGE;
        }
        // ================= STAGE: set_phi + predict_stage_sdirk (runge_kutta.rs:645-665) ========================
        if (__any_sync(0xffffffffu, state == R_STAGE) && state == R_STAGE) {
            const int i = stage;
            t_stage = t + pa.rk.c[i] * h;
            double ph[N];
#pragma unroll
            for (int k = 0; k < N; ++k) ph[k] = SY(k);
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
                const double aij = pa.rk.a[j * ns + i];
#pragma unroll
                for (int k = 0; k < N; ++k) ph[k] = SDF(j, k) * aij + ph[k];
            }
#pragma unroll
            for (int k = 0; k < N; ++k) SPHI(k) = ph[k];
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < N; ++k) x_cur[k] = h * SDY(k);
            } else if (i == 1) {
#pragma unroll
                for (int k = 0; k < N; ++k) x_cur[k] = SDF(0, k);
            } else {
                const double cc = DSB_DIV(pa.rk.c[i] - pa.rk.c[i - 2], pa.rk.c[i - 1] - pa.rk.c[i - 2]);
                const double al = -cc, be = 1.0 + cc;
#pragma unroll
                for (int k = 0; k < N; ++k) x_cur[k] = al * SDF(i - 2, k) + be * SDF(i - 1, k);
            }
            conv.reset();
            if (!is_jacobian_set) { jac_kind = DSB_KIND_LAZY; after_jac = R_NEWTON; state = R_JAC; }
            else state = R_NEWTON;
        }

        // ================= NEWTON: one iteration on F(x) = M x - h f(phi + c x) =================================
        // (up to newton_passes iterations per trip, rolled: see dsb_bdf_kernel.cuh)
#pragma unroll 1
        for (int pass = 0; pass < newton_passes; ++pass) {
        if (!__any_sync(0xffffffffu, state == R_NEWTON)) break;
        if (state == R_NEWTON) {
            double pl[NP > 0 ? NP : 1];
#pragma unroll
            for (int j = 0; j < NP; ++j) pl[j] = SP(j);
            double delta[N];
            bool sens_eq = false;
            if constexpr (SENS) sens_eq = eq > 0;
            {
                double tmpv[N];
                if (sens_eq) {
                    if constexpr (SENS) {
                        // SdirkCallable<SensEquations>::call_inplace on SensRhs::call_inplace: J(stage value) (phi_s + c x) + f_p e_q
                        double ysv[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) { tmpv[i] = cg * x_cur[i] + SPHS(i); ysv[i] = SOY(i); }
                        M::jac_mul(ysv, pl, t_stage, tmpv, delta);
                        st.v[DSB_STAT_RHS_JAC_MULS] += 1;
#pragma unroll
                        for (int i = 0; i < N; ++i) delta[i] += SFP(eq - 1, i);
                    }
                } else {
#pragma unroll
                for (int i = 0; i < N; ++i) tmpv[i] = cg * x_cur[i] + SPHI(i);
                M::rhs(tmpv, pl, t_stage, delta);
                st.v[DSB_STAT_RHS_CALLS] += 1;
                }
                const double beta = -op_h;
                if (M::HAS_MASS) {
                    M::mass(x_cur, pl, t_stage, beta, delta);
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) delta[i] = x_cur[i] + beta * delta[i];
                }
            }
            LaneLU<N> lu;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                lu.piv[j] = (int)((piv_packed >> (4 * j)) & 15ull);
#pragma unroll
                for (int i = 0; i < N; ++i) lu.a[j][i] = SLU(j, i);
            }
            if (!lu.solve(delta)) {
                newton_ok = false; state = R_POST;
            } else {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    x_cur[i] -= delta[i];
                    double w = wt[i];
                    if constexpr (SENS) { if (sens_eq) w = dsb_abs(SSS(eq - 1, i)) * pa.rtol + pa.atol[i]; }   // error_y = state.s[q]
                    const double term = DSB_DIV(delta[i], w);
                    acc += term * term;
                }
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
            st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;      // (sensitivity solves too, failed ones included)
            bool stage_done = true;
            if constexpr (SENS) {
                if (newton_ok) {
                    const int i = stage;
                    if (eq == 0) {
                        double ysv[N];
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            ysv[k] = cg * x_cur[k] + SPHI(k);
                            SOY(k) = ysv[k];
                            SDF(i, k) = x_cur[k];
                        }
                        sens_update_state(ysv, t_stage);            // f_p at the stage value (runge_kutta.rs:693-695)
                    } else {
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            SOS(eq - 1, k) = cg * x_cur[k] + SPHS(k);
                            SSD(eq - 1, i, k) = x_cur[k];
                        }
                    }
                    if (eq < NP) {
                        sens_stage_setup(eq);
                        eq += 1;
                        state = R_NEWTON;
                        stage_done = false;
                    } else {
                        eq = 0;
                        stage = i + 1;
                        state = (stage < ns) ? R_STAGE : R_ERRTEST;
                        stage_done = false;
                    }
                } else {
                    eq = 0;
                }
            }
            if (!stage_done) {
                // handled above
            } else if (newton_ok) {
                const int i = stage;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    SOY(k) = cg * x_cur[k] + SPHI(k);          // get_f_eval: stage value
                    SDF(i, k) = x_cur[k];
                }
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
#undef SM
#undef DSB_DIV
#undef SDF
#undef SJ
#undef SMM
#undef SLU
#undef SY
#undef SDY
#undef SOY
#undef SPHI
#undef SP
#undef SSD
#undef SSS
#undef SDS
#undef SOS
#undef SFP
#undef SPHS
}
