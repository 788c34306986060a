// dsb_bdf_kernel.cuh -- `problem.bdf::<LS>()?.solve_dense(t_eval)` for every instance of a batch.
//
// Execution model (B200-first, not the reference's): one CUDA thread ("lane") integrates one instance
// at a time, with its OWN step size, order, Newton state and counters.  The reference's nested loops
// (step -> retry loop -> Newton loop) are flattened into a per-lane STATE MACHINE whose blocks each
// appear exactly once in the kernel:
//
//     FETCH -> JAC(construct) -> TSTOP(first) -> PREDICT -> NEWTON* -> POST -> [SELECT] -> [RESCALE -> JAC]
//           -> TSTOP -> OUTPUT -> PREDICT -> ...                  \-> (retry) JAC / RESCALE -> PREDICT
//
// In the loop body the rarely needed blocks come first and POST last, so that a lane flows
// [SELECT -> RESCALE -> JAC ->] TSTOP -> OUTPUT -> PREDICT -> NEWTON -> POST in ONE trip and the blocks
// every trip needs are contiguous in the instruction cache.
//
// so that (a) the 32 lanes of a warp, which are at different points of their own integrations, still
// execute the hot block (one Newton iteration) together instead of serialising whole retry loops,
// (b) the code stays small enough for the instruction cache (the nested-loop version inlined the
// rescale / refactor blocks at every call site: 13 k SASS instructions, 70 % of issue slots stalled on
// instruction fetch, ncu profiles/r1_v1_*), and (c) a lane that finishes its instance fetches the next
// one from a global work counter, so a warp never waits for its slowest instance.
//
// Storage: Newton work vectors and controller scalars in registers; the difference array D (n x 8),
// df/dy, M (DAEs), the LU factors, state.y, the predictor and the parameters in shared memory, one
// column per thread (word w of thread t at smem[w * blockDim + t]: conflict-free, indexable by the
// run-time order).  Global memory is touched only to fetch an instance and to write its outputs.
//
// Arithmetic: every floating-point expression keeps the operation order of the reference's CPU path so
// that a build with --fmad=false reproduces its controller decisions bit for bit (checked against the
// oracle in tests/).  Restated functions (paths relative to /root/reference/crates/diffsol/src):
//   Bdf::_new                    ode_solver/bdf.rs:230-368
//   Bdf::step                    ode_solver/bdf.rs:1277-1589   (retry loop, order/step selection)
//   _predict_forward / set_psi   ode_solver/bdf.rs:667-692, op/bdf.rs:182-210
//   BdfCallable::call_inplace    op/bdf.rs:240-256             F(y) = M (y - y0 + psi) - c f(y)
//   BdfCallable::jacobian_inplace op/bdf.rs:273-300            A = M - c J
//   newton_iteration + NoLineSearch  crates/diffsol-nl/src/newton.rs:13-36, line_search.rs:48-69
//   Convergence                  crates/diffsol-nl/src/convergence.rs:64-139
//   _jacobian_updates            ode_solver/bdf.rs:465-506, jacobian_update.rs
//   _update_step_size, _compute_r ode_solver/bdf.rs:433-463, 508-577
//   _update_diff                 ode_solver/bdf.rs:646-664
//   error_control, predict_error_control   ode_solver/bdf.rs:812-932
//   pi_controller_raw            ode_solver/runge_kutta.rs:1313-1335
//   handle_tstop, set_stop_time  ode_solver/bdf.rs:694-731, 1591-1599
//   interpolate                  ode_solver/bdf.rs:767-782, 1080-1106
//   fn solve_dense               ode_solver/method.rs:721-848
#pragma once
#include "dsb_lane.cuh"
#include "dsb_init_kernel.cuh"      // lane_consistent_solve: consistent sensitivities of a DAE
#include "dsb_roots.cuh"

#define DSB_NSTATS_USED 13      // counters the kernels maintain (the C ABI rows have DSB_NSTATS = 16 slots)
enum dsb_lane_state {
    L_FETCH = 0, L_POST, L_SELECT, L_RESCALE, L_JAC, L_TSTOP, L_OUTPUT, L_PREDICT, L_NEWTON, L_FINISH, L_IDLE,
    L_REINIT            // equations with a reset function: the is_state_modified branch of Bdf::step after a reset
};
#define DSB_KIND_CONSTRUCT 5      // Bdf::_new's reset_jacobian (counted as a checkpoint setup, bdf.rs:351-359)

template <class M>
struct BdfLayout {
    static constexpr int N = M::N, NP = M::NP;
    static constexpr int O_D = 0;                                   // D[DSB_NDIFF][N]
    static constexpr int O_J = O_D + DSB_NDIFF * N;                 // rhs_jac[col][row]
    static constexpr int O_M = O_J + N * N;                         // mass_jac[col][row] (DAE only)
    static constexpr int O_LU = O_M + (M::HAS_MASS ? N * N : 0);    // LU factors [col][row]
#ifdef DSB_LANE_LU_RCP
    static constexpr int O_RCP = O_LU + N * N;                      // RN(1 / U_ii) for the back substitution (dsb_div_rcp)
    static constexpr int O_Y = O_RCP + N;                           // state.y
#else
    static constexpr int O_Y = O_LU + N * N;                        // state.y
#endif
    static constexpr int O_YP = O_Y + N;                            // y_predict
    static constexpr int O_P = O_YP + N;                            // parameters
    static constexpr int O_ST = O_P + (NP > 0 ? NP : 1);            // statistics, two int32 per word
    // forward sensitivities (DsbWithSens<M> only): per parameter a difference array, state.s, s_delta and the column of
    // f_p at the predictor; the predictor of the sensitivity being solved; the parked solution of the main Newton solve.
    // (Tried and rejected: the difference arrays sdiff in a lane-interleaved global-memory slot, touched once per step --
    // 166 -> 94 words of shared memory per lane, 5 -> 9 warps per SM for Robertson -- 760 -> 1166 ms per 10^6 instances:
    // the blocks that touch sdiff run with ~3 active lanes per warp, and their L2 round trips are not hidden.)
    static constexpr bool SENS = dsb_model_sens_on<M>::value;
    // SDIFF_GLOBAL: the difference arrays of the sensitivities live in a lane-interleaved global-memory slot (L2-resident:
    // 148 blocks x THREADS x NP x 8 x N words) and every predictor / psi pair is formed in PREDICT, where the lanes arrive in
    // groups (end-of-step hold, slow pool), and kept in shared memory for the NP solves of the attempt.
#ifdef DSB_SENS_SDIFF_GLOBAL
    static constexpr bool SDIFF_GLOBAL = SENS;
#else
    static constexpr bool SDIFF_GLOBAL = false;
#endif
    static constexpr int SDIFF_WORDS = NP * DSB_NDIFF * N;          // per lane, shared or global
    static constexpr int O_SDF = O_ST + (DSB_NSTATS_USED + 1) / 2;  // sdiff[NP][DSB_NDIFF][N]
    static constexpr int O_SS = O_SDF + (SDIFF_GLOBAL ? 0 : SDIFF_WORDS);   // state.s[NP][N]
    static constexpr int O_SDL = O_SS + NP * N;                     // s_deltas[NP][N]
    static constexpr int O_SFP = O_SDL + NP * N;                    // f_p e_q at (y_predict, t_predict) [NP][N]
    static constexpr int O_SPR = O_SFP + NP * N;                    // s_predict[N] (SDIFF_GLOBAL: [NP][N])
    static constexpr int O_SPS = O_SPR + (SDIFF_GLOBAL ? NP * N : N);       // SDIFF_GLOBAL: psi - s_predict of every sensitivity [NP][N]
    static constexpr int O_SYC = O_SPS + (SDIFF_GLOBAL ? NP * N : 0);       // the main solve's y while the sensitivities are solved
    static constexpr int WORDS = SENS ? O_SYC + N : O_SDF;
    static constexpr int THREADS = LaneBlockShape<WORDS, N>::THREADS;
    static constexpr int MAXNREG = LaneBlockShape<WORDS, N>::MAXNREG;
};

template <class M>
__global__ void __maxnreg__(BdfLayout<M>::MAXNREG) dsb_bdf_solve_dense_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                                              const __grid_constant__ DsbBatchBuffers bb,
                                                                              unsigned long long* __restrict__ work_counter) {
    constexpr int N = M::N;
    constexpr int NP = M::NP;
    static_assert(N <= 16, "pivots are packed 4 bits per row");
    typedef BdfLayout<M> Lay;
    extern __shared__ double dsb_lane_smem[];
    double* const sm = dsb_lane_smem + threadIdx.x;
#define SM(w) sm[(w) * Lay::THREADS]
#define DSB_DIV(a, b) DsbDivShared::div((a), (b))      // one shared division routine (code size, dsb_math.h)
    // x / N (the mean of a weighted norm's squared terms) through the compile-time RN(1 / N): the same quotient in 5
    // operations (dsb_math.h: dsb_div_rcp)
    // (tried and rejected: x / N and the x / i of the rescale rows through dsb_div_rcp with constant reciprocals -- fewer
    // operations, but ten more inline expansions: 59.4 -> 67.9 ms, the loop body no longer fits the instruction cache)
#ifdef DSB_OPT_DIVN
#define DSB_DIV_N(x) dsb_div_rcp((x), (double)N, 1.0 / (double)N)
#else
#define DSB_DIV_N(x) DSB_DIV((x), (double)N)
#endif
#ifndef DSB_NEWTON_DIV
#define DSB_NEWTON_DIV DsbDivShared               // tuning experiment: inline expansion at the six hottest sites
#endif
#define SD(j, i) SM(Lay::O_D + (j) * N + (i))
#define SJ(j, i) SM(Lay::O_J + (j) * N + (i))
#define SMM(j, i) SM(Lay::O_M + (j) * N + (i))
#define SLU(j, i) SM(Lay::O_LU + (j) * N + (i))
#define SY(i) SM(Lay::O_Y + (i))
#define SYP(i) SM(Lay::O_YP + (i))
#define SP(i) SM(Lay::O_P + (i))
#define SDF(q, j, i) (*(Lay::SDIFF_GLOBAL ? sdf_g + (size_t)(((q) * DSB_NDIFF + (j)) * N + (i)) * sdf_stride : &SM(Lay::O_SDF + ((q) * DSB_NDIFF + (j)) * N + (i))))
#define SSS(q, i) SM(Lay::O_SS + (q) * N + (i))
#define SDL(q, i) SM(Lay::O_SDL + (q) * N + (i))
#define SFP(q, i) SM(Lay::O_SFP + (q) * N + (i))
#define SPR(q, i) SM(Lay::O_SPR + (Lay::SDIFF_GLOBAL ? (q) * N : 0) + (i))
#define SPS(q, i) SM(Lay::O_SPS + (q) * N + (i))
#define SYC(i) SM(Lay::O_SYC + (i))
    constexpr bool SENS = Lay::SENS;
    const size_t sdf_stride = (size_t)gridDim.x * Lay::THREADS;
    double* const sdf_g = Lay::SDIFF_GLOBAL ? bb.sens_ws + ((size_t)blockIdx.x * Lay::THREADS + threadIdx.x) : nullptr;
    static_assert(!SENS || (dsb_model_has_sens<M>::value && dsb_model_nroots<M>::value == 0 &&
                            !dsb_model_nout<M>::has_out && !dsb_model_has_reset<M>::value),
                  "sensitivities: equations with sens_mul / init_sens, no root / output / reset functions");

    const int64_t B = pa.nbatch;
    const int nt = pa.nt;
    const bool free_running = pa.free_running != 0;
    const int quorum = pa.quorum;
    const int newton_passes = pa.newton_passes < 1 ? 1 : pa.newton_passes;
    const double eps = 2.220446049250313e-16;

    // ---- per-lane registers -----------------------------------------------------------------------
    int state = L_FETCH;
    int64_t inst = 0;
    // BdfState / Bdf scalars
    int order = 1, n_equal_steps = 0;
    double t = 0.0, h = 0.0, c = 0.0, t_predict = 0.0;
    bool has_tstop = false, has_prev_error = false, jacobian_is_stale = true;
    double tstop = 0.0, prev_error_norm = 0.0;
    LaneJacobianUpdate ju; ju.init(1.0);
    LaneConvergence conv;
    conv.tol = pa.opt.nonlinear_solver_tolerance; conv.max_iter = pa.opt.max_nonlinear_solver_iterations;
    conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
    SmemLaneStats<2 * Lay::THREADS> st;
    st.v.base = reinterpret_cast<int*>(&SM(Lay::O_ST));
    unsigned long long piv_packed = 0;         // 4 bits per row
    // Newton work vectors
    double y_cur[N], psi_neg_y0[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { y_cur[i] = 0.0; psi_neg_y0[i] = 0.0; }
    // step()-local state
    bool convergence_fail = false, newton_ok = false, first = true, reached = false, accepted = false;
    bool repredict = true, pending_etf = false, rs_ignore_small = false;
    int old_num_error_test_failures = 0, col = 0;
    double safety = 0.0, error_norm = 0.0;
    // pending control transfers
    int after_rescale = L_JAC, after_jac = L_TSTOP, jac_kind = DSB_CHECKPOINT;
    double rescale_factor = 1.0;
    // sensitivities: which equation the Newton block is solving (0 = the state, q + 1 = the sensitivity to parameter q)
    // and the sensitivity residual's OWN c, which stays 0 until the first step-size update (op/bdf.rs:61, bdf.rs:551-553)
    int eq = 0;
    double c_sens = 0.0;

    int fin_status = DSB_STATUS_OK;
    auto finish = [&](int status) { fin_status = status; state = L_FINISH; };
    // root finding (nonlinear_solver/root.rs; bdf.rs:143, 301-306, 1566-1579): only compiled for equations with roots
    constexpr int NR = dsb_model_nroots<M>::value;
    LaneRootFinder<(NR > 0 ? NR : 1), DsbDivShared> rf;
    rf.t0 = 0.0;
    int root_found = -1;
#pragma unroll
    for (int r = 0; r < (NR > 0 ? NR : 1); ++r) rf.g0[r] = 0.0;
    // interpolate (bdf.rs:767-782, 1080-1106) at tq <= t from the difference array
    auto interpolate = [&](double tq, double (&yo)[N]) {
        double time_factor = 1.0;
#pragma unroll
        for (int i = 0; i < N; ++i) yo[i] = SD(0, i);
#pragma unroll 1
        for (int j = 0; j < order; ++j) {
            const double j_t = (double)j;
            time_factor *= DSB_DIV(tq - (t - h * j_t), h * (1.0 + j_t));
#pragma unroll
            for (int i = 0; i < N; ++i) yo[i] = time_factor * SD(j + 1, i) + yo[i];
        }
    };
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
    // bdf.rs:694-731.  0 = nothing, 1 = TstopReached, 2 = step size must be clipped (rescale_factor set), < 0 = -status
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
#ifndef DSB_NO_HOST_TABLES     // A/B switch: 61.5 -> 59.4 ms per 10^6 Robertson instances (profiles/r2_lane_kernel_ab.log)
        const double ki = pa.tab.pi_ki[eff_order];                 // pi_control_integral / order, divided on the host
        const bool p_only = pa.opt.pi_control_proportional == 0.0 || !has_prev_error;
        const double kp = p_only ? 0.0 : pa.tab.pi_kp[eff_order];
#else
        const double order_f = (double)eff_order;
        const double ki = DSB_DIV(pa.opt.pi_control_integral, order_f);
        const bool p_only = pa.opt.pi_control_proportional == 0.0 || !has_prev_error;
        const double kp = p_only ? 0.0 : DSB_DIV(pa.opt.pi_control_proportional, order_f);
#endif
        double v = dsb_pow(err, p_only ? -ki : -(ki + kp));
        if (!p_only) v = v * dsb_pow(prev_error_norm, kp);
        return v;
    };
    // ||D[:, j]||^2_w(state.y) (vector/nalgebra_serial.rs:395-408)
    // x_i / (|ref_i| rtol + atol_i), i = 0 .. N-1, squared and summed in index order (vector/nalgebra_serial.rs:395-408)
    auto weighted_sum = [&](const double (&x)[N], const double (&ref)[N]) -> double {
#ifdef DSB_DIV_VEC
        DsbVecN<N> a, b;
#pragma unroll
        for (int i = 0; i < N; ++i) { a.v[i] = x[i]; b.v[i] = dsb_abs(ref[i]) * pa.rtol + pa.atol[i]; }
        const DsbVecN<N> q = dsb_div_vec_fn<N>(a, b);
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) acc += q.v[i] * q.v[i];
        return acc;
#else
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double term = DSB_DIV(x[i], dsb_abs(ref[i]) * pa.rtol + pa.atol[i]);
            acc += term * term;
        }
        return acc;
#endif
    };
    auto diff_col_norm = [&](int j) -> double {
        double x[N], ref[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { x[i] = SD(j, i); ref[i] = SY(i); }
        return DSB_DIV_N(weighted_sum(x, ref));
    };
    // ---- sensitivities (only instantiated for DsbWithSens<M>) ----
    // squared_norm(x, ref, sens_atol, sens_rtol) (bdf.rs:844-858, 908-919)
    auto sens_weighted_norm = [&](const double (&x)[N], const double (&ref)[N]) -> double {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double term = DSB_DIV(x[i], dsb_abs(ref[i]) * pa.sens_rtol + pa.sens_atol[i]);
            acc += term * term;
        }
        return DSB_DIV_N(acc);
    };
    // f_p e_q at (x, tq) for every parameter: SensRhs::update_state through _default_sens_inplace
    // (sens_equations.rs:129-134, op/nonlinear_op.rs:72-81)
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
    // the start of the Newton solve for sensitivity q (bdf.rs:948-968): predictor and psi from sdiff[q], s_new <- s_predict
    auto sens_setup = [&](int q) {
        if constexpr (Lay::SDIFF_GLOBAL) {          // formed by sens_predict_all() in PREDICT
#pragma unroll
            for (int i = 0; i < N; ++i) { psi_neg_y0[i] = SPS(q, i); y_cur[i] = SPR(q, i); }
            conv.reset();
            return;
        }
        // (one pass over the columns: each accumulator sees its terms in the reference's order)
        double sp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { const double d0 = SDF(q, 0, i); sp[i] = 0.0; sp[i] += d0; }
#pragma unroll
        for (int i = 0; i < N; ++i) { const double d1 = SDF(q, 1, i); sp[i] += d1; psi_neg_y0[i] = pa.tab.gamma[1] * d1; }
#pragma unroll 1
        for (int j = 2; j <= order; ++j) {
            const double g = pa.tab.gamma[j];
#pragma unroll
            for (int i = 0; i < N; ++i) { const double dj = SDF(q, j, i); sp[i] += dj; psi_neg_y0[i] = g * dj + psi_neg_y0[i]; }
        }
        const double a = pa.tab.alpha[order];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            psi_neg_y0[i] *= a;
            psi_neg_y0[i] -= sp[i];
            SPR(q, i) = sp[i];
            y_cur[i] = sp[i];
        }
        conv.reset();
    };
    // SDIFF_GLOBAL: the same sums for every sensitivity at once, the column loop unrolled under a predicate so that the
    // loads of all columns are in flight together (each accumulator still sees its terms in the reference's order)
    auto sens_predict_all = [&]() {
        if constexpr (Lay::SDIFF_GLOBAL) {
            double sp[NP > 0 ? NP : 1][N], ps[NP > 0 ? NP : 1][N];
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
                for (int i = 0; i < N; ++i) { const double d0 = SDF(q, 0, i); sp[q][i] = 0.0; sp[q][i] += d0; }
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
                for (int i = 0; i < N; ++i) { const double d1 = SDF(q, 1, i); sp[q][i] += d1; ps[q][i] = pa.tab.gamma[1] * d1; }
#pragma unroll
            for (int j = 2; j <= DSB_MAX_ORDER; ++j) {
                if (j <= order) {
                    const double g = pa.tab.gamma[j];
#pragma unroll
                    for (int q = 0; q < NP; ++q)
#pragma unroll
                        for (int i = 0; i < N; ++i) { const double dj = SDF(q, j, i); sp[q][i] += dj; ps[q][i] = g * dj + ps[q][i]; }
                }
            }
            const double a = pa.tab.alpha[order];
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    ps[q][i] *= a;
                    ps[q][i] -= sp[q][i];
                    SPR(q, i) = sp[q][i];
                    SPS(q, i) = ps[q][i];
                }
        }
    };

#ifdef DSB_LANE_PROFILE          // warp-scheduler occupancy counters (warp-uniform values, lane 0 publishes them)
    unsigned long long prof[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) prof[k] = 0;
#define DSB_PROF_BLOCK(k, cond) { const int n_ = __popc(__ballot_sync(0xffffffffu, (cond))); if (n_) { prof[k] += 1; prof[k + 1] += n_; } }
#else
#define DSB_PROF_BLOCK(k, cond)
#endif
    while (true) {
        // ---- warp-level block scheduler ---------------------------------------------------------------
        // The lanes of a warp are in different states.  The blocks every lane passes through once per
        // Newton iteration or step (POST, TSTOP, OUTPUT, PREDICT, NEWTON) run whenever a lane needs them;
        // a lane flows through all of them within one trip of this loop.  The heavy blocks only a few lanes
        // need at a time (SELECT: order/step selection with its three pow()s; RESCALE + JAC: step-size
        // change and refactorisation) would run with a handful of active lanes on every trip, so a lane
        // that needs one WAITS until pa.quorum lanes (or half of the active ones) want the same group.
        // Waiting never changes a lane's arithmetic, only when it runs.
#ifdef DSB_BLOCK_SYNC           // experiment: the warps of a block walk the loop body together (instruction-cache locality)
        if (__syncthreads_and(state == L_IDLE)) break;
        const unsigned m_idle = __ballot_sync(0xffffffffu, state == L_IDLE);
        if (m_idle == 0xffffffffu) continue;
#else
        const unsigned m_idle = __ballot_sync(0xffffffffu, state == L_IDLE);
        if (m_idle == 0xffffffffu) break;
#endif
        const int n_active = 32 - __popc(m_idle);
        const int n_slow = __popc(__ballot_sync(0xffffffffu, state == L_SELECT || state == L_RESCALE || state == L_JAC));
        // one pool: SELECT -> RESCALE -> JAC flow through together (separate pools and a quorum on POST were measured
        // and rejected)
        const bool run_select = n_slow > 0 && (n_slow >= quorum || 2 * n_slow >= n_active);
        const bool run_setup = run_select;
#ifdef DSB_LANE_PROFILE
        prof[0] += 1; prof[1] += n_active; prof[2] += (run_select ? 0 : n_slow);
        DSB_PROF_BLOCK(3, run_select && (state == L_SELECT || state == L_RESCALE || state == L_JAC))
#endif

        // ================= FINISH: write the instance's results, then fetch the next one =====================
        if (__any_sync(0xffffffffu, state == L_FINISH) && state == L_FINISH) {
            bb.status[inst] = fin_status;
            bb.fin_t[inst] = t; bb.fin_h[inst] = h; bb.fin_order[inst] = order;
            bb.ncols[inst] = col;
            if (NR > 0) bb.root_idx[inst] = root_found;
#pragma unroll
            for (int k = 0; k < DSB_NSTATS_USED; ++k) bb.stats[(int64_t)k * B + inst] = st.v[k];
            state = L_FETCH;
        }
        // ================= FETCH: next instance from the work counter; Bdf::_new part 1 ==================
        if (__any_sync(0xffffffffu, state == L_FETCH) && state == L_FETCH) {
            inst = (int64_t)atomicAdd(work_counter, 1ull);
            if (inst >= B) {
                state = L_IDLE;
            } else if (bb.status[inst] == DSB_STATUS_OK) {      // else: initialisation failed, keep its status
#pragma unroll
                for (int j = 0; j < NP; ++j) SP(j) = bb.params[(int64_t)j * B + inst];
#pragma unroll
                for (int k = 0; k < DSB_NSTATS_USED; ++k) st.v[k] = bb.stats[(int64_t)k * B + inst];
                order = 1; n_equal_steps = 0;
                t = pa.t0; h = bb.h0[inst];
                conv.eta = pa.tab.eta_reset; conv.old_norm = 0.0; conv.reset();
#pragma unroll
                for (int j = 0; j < DSB_NDIFF; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) SD(j, i) = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double yi = bb.y0[(int64_t)i * B + inst];
                    SY(i) = yi; SD(0, i) = yi; SD(1, i) = bb.dy0[(int64_t)i * B + inst] * h;
                }
                if constexpr (NR > 0) {                         // Bdf::_new: root_finder.init(root_fn, state.y, state.t)
                    double y0l[N], pl0[NP > 0 ? NP : 1];
#pragma unroll
                    for (int i = 0; i < N; ++i) y0l[i] = SY(i);
#pragma unroll
                    for (int j = 0; j < NP; ++j) pl0[j] = SP(j);
                    M::root(y0l, pl0, t, rf.g0);
                    rf.t0 = t; root_found = -1;
                }
                if constexpr (SENS) {
                    // state.rs:1158-1180 (s_q = (d y0 / d p) e_q), :178-189 (ds_q = J(y0) s_q + f_p e_q), bdf_state.rs:88-98
                    // (sdiff[q][:, 0] = s_q, [:, 1] = h ds_q)
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
                        for (int i = 0; i < N; ++i) {
                            dsq[i] += SFP(q, i);
                            SSS(q, i) = sq[i]; SDL(q, i) = M::HAS_MASS ? dsq[i] : 0.0;      // (a DAE's ds_q is parked for the solve below)
                            SDF(q, 0, i) = sq[i]; SDF(q, 1, i) = dsq[i] * h;
                        }
#pragma unroll 1
                        for (int j = 2; j < DSB_NDIFF; ++j)
#pragma unroll
                            for (int i = 0; i < N; ++i) SDF(q, j, i) = 0.0;
                    }
                    if constexpr (M::HAS_MASS) {
                        // set_consistent_augmented's algebraic part (state.rs:191-237): after every ds_q is formed, each
                        // sensitivity vector goes through the InitOp solve on SensRhs (J(y0) x + f_p e_q; its Jacobian is
                        // df/dy at y0, evaluated per parameter), with ONE Convergence for all of them
                        LaneConvergence ic_conv;
                        ic_conv.tol = pa.opt.nonlinear_solver_tolerance;
                        ic_conv.eta = pa.tab.eta_reset;
                        ic_conv.max_iter = pa.opt.ic_max_newton_iterations;
                        ic_conv.reset();
                        int ic_status = DSB_STATUS_OK;
#pragma unroll 1
                        for (int q = 0; q < NP && ic_status == DSB_STATUS_OK; ++q) {
#pragma unroll
                            for (int i = 0; i < N; ++i) { sq[i] = SSS(q, i); dsq[i] = SDL(q, i); SDL(q, i) = 0.0; }
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
                            for (int i = 0; i < N; ++i) { SSS(q, i) = sq[i]; SDF(q, 0, i) = sq[i]; SDF(q, 1, i) = dsq[i] * h; }
                        }
                        if (ic_status != DSB_STATUS_OK) finish(ic_status);
                    }
                    eq = 0; c_sens = 0.0;
                }
                c = h * pa.tab.alpha[1];
                jacobian_is_stale = true;
                ju.init(1.0);                                   // jacobian_update.rs:27 -- h_at_last starts at ONE
                has_tstop = false; tstop = 0.0; has_prev_error = false; prev_error_norm = 0.0;
                convergence_fail = false; first = true; reached = false; pending_etf = false; col = 0;
                t_predict = t;
                jac_kind = DSB_KIND_CONSTRUCT;
                if constexpr (SENS && M::HAS_MASS) { if (state != L_FINISH) state = L_JAC; }      // (the sensitivities' consistency solve may fail)
                else state = L_JAC;
            }
        }
        // ================= REINIT: Bdf::step finds the state modified by a reset (bdf.rs:1291-1318) ==========
        // root finder re-initialised, difference array back to first order (D[:, 0] = y, D[:, 1] = h dy: the dy of
        // apply_reset is parked in the predictor's words; the higher columns keep what they held), _jacobian_updates(c,
        // StepSuccess), then set_stop_time again: the TSTOP block's first-step path.  Only compiled for such equations.
        if constexpr (dsb_model_has_reset<M>::value) {
            if (__any_sync(0xffffffffu, state == L_REINIT) && state == L_REINIT) {
                double yl[N], pl[NP > 0 ? NP : 1];
#pragma unroll
                for (int i = 0; i < N; ++i) yl[i] = SY(i);
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = SP(j);
                M::root(yl, pl, t, rf.g0);
                rf.t0 = t;
                order = 1; n_equal_steps = 0;
#pragma unroll
                for (int i = 0; i < N; ++i) { SD(0, i) = yl[i]; SD(1, i) = SYP(i) * h; }
                c = h * pa.tab.alpha[1];
                has_prev_error = false;
                jac_kind = DSB_STEP_SUCCESS; after_jac = L_TSTOP;
                first = true;
                state = L_JAC;
            }
        }
        // ================= SELECT: order / step-size selection after an accepted step (bdf.rs:1489-1563), ==
        // or the shrink factor after a failed error test (bdf.rs:1431-1442).  Every pow() of the controller
        // is issued from ONE call site inside a rolled loop so that the lanes of the warp share it.
        if (run_select && state == L_SELECT) {
            const int ord = order;
            const double inf = dsb_from_bits(0x7ff0000000000000ULL);
            // the two neighbouring-order error estimates and the three controller factors come out of ONE rolled
            // loop, so that the weighted norm and pow() each have a single call site (code size)
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll 1
            for (int q = 0; q < 3; ++q) {
                if (accepted || q == 1) {
                    double err = error_norm;
                    if (q != 1) {
                        err = inf;
                        if ((q == 0) ? (ord > 1) : (ord < DSB_MAX_ORDER)) {
                            const double e = diff_col_norm(ord + q) * pa.tab.error_const2[ord - 1 + q];
                            err = (0.0 < e) ? e : 0.0;
                            if constexpr (SENS) {               // predict_error_control's sensitivity terms (bdf.rs:908-919)
                                if (pa.sens_error_control) {
                                    double xa[Lay::SDIFF_GLOBAL ? (NP > 0 ? NP : 1) : 1][N];
                                    if constexpr (Lay::SDIFF_GLOBAL) {      // the NP columns' loads in flight together (global slot)
#pragma unroll
                                        for (int qs = 0; qs < NP; ++qs)
#pragma unroll
                                            for (int i = 0; i < N; ++i) xa[qs][i] = SDF(qs, ord + q, i);
                                    }
#pragma unroll 1
                                    for (int qs = 0; qs < NP; ++qs) {
                                        double x[N], ref[N];
                                        if constexpr (Lay::SDIFF_GLOBAL) {
                                            dsb_static_for<0, (NP > 0 ? NP : 1)>([&](auto Q) {
                                                constexpr int QC = decltype(Q)::value;
                                                if (QC == qs) {
#pragma unroll
                                                    for (int i = 0; i < N; ++i) x[i] = xa[QC][i];
                                                }
                                            });
                                        } else {
#pragma unroll
                                            for (int i = 0; i < N; ++i) x[i] = SDF(qs, ord + q, i);
                                        }
#pragma unroll
                                        for (int i = 0; i < N; ++i) ref[i] = SSS(qs, i);
                                        const double es = sens_weighted_norm(x, ref) * pa.tab.error_const2[ord - 1 + q];
                                        err = (err < es) ? es : err;
                                    }
                                }
                            }
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

        // ================= RESCALE: _update_step_size(factor) (bdf.rs:508-577) ============================
        // D[:, 0..=k] <- D[:, 0..=k] * (R(k, factor) * U(k)).  R and RU are produced one ROW at a time and
        // folded into the new columns at once.  Terms that are exactly zero in the reference's gemm
        // (U is upper triangular with U[l, j] = (-1)^l C(j, l); R[i, 0] = RU[i, 0] = RU[0, i] = delta_i0, so
        // column 0 is unchanged) are skipped: adding an exact zero never changes a non-zero partial sum.
        if (run_setup && state == L_RESCALE) {
            const double factor = rescale_factor;
            const double new_h = factor * h;
            n_equal_steps = 0;
            const int k = order;
            const double* __restrict__ u = pa.tab.u[DSB_MAX_ORDER];         // leading dimension 6
            double rrow[DSB_MAX_ORDER + 1];
            double nd[DSB_MAX_ORDER + 1][N];
            // The row loop is ROLLED (code size: the unrolled version was 720 SASS instructions, a fifth of the
            // kernel, and the loop body of the kernel did not fit the 32 KB instruction cache).  Accumulators
            // start at -0.0: (-0.0) + x == x bit for bit for every x, so the first term is "assigned" as in
            // nalgebra's gemm with beta = 0.  x / 1, x / 2, x / 4 are exact, so `num / i` covers every row.
#pragma unroll
            for (int l = 1; l <= DSB_MAX_ORDER; ++l) {
                rrow[l] = 1.0;
#pragma unroll
                for (int s = 0; s < N; ++s) nd[l][s] = -0.0;
            }
#pragma unroll 1
            for (int i = 1; i <= k; ++i) {
                const double i_t = (double)i;
#pragma unroll
#ifdef DSB_OPT_RESCALE_RCP
                for (int l = 1; l <= DSB_MAX_ORDER; ++l) rrow[l] = dsb_div_rcp(rrow[l] * (i_t - 1.0 - factor * (double)l), i_t, pa.tab.inv_int[i]);
#else
                for (int l = 1; l <= DSB_MAX_ORDER; ++l) rrow[l] = DSB_DIV(rrow[l] * (i_t - 1.0 - factor * (double)l), i_t);
#endif
                double di[N];
#pragma unroll
                for (int s = 0; s < N; ++s) di[s] = SD(i, s);
#pragma unroll
                for (int j = 1; j <= DSB_MAX_ORDER; ++j) {
                    double ru_ij = rrow[1] * u[j * 6 + 1];              // RU[i, j] = sum_{l <= j} R[i, l] U[l, j]
#pragma unroll
                    for (int l = 2; l <= j; ++l) ru_ij = rrow[l] * u[j * 6 + l] + ru_ij;
#pragma unroll
                    for (int s = 0; s < N; ++s) nd[j][s] = di[s] * ru_ij + nd[j][s];
                }
            }
#pragma unroll
            for (int j = 1; j <= DSB_MAX_ORDER; ++j) {
                if (j <= k) {
#pragma unroll
                    for (int s = 0; s < N; ++s) SD(j, s) = nd[j][s];
                }
            }
            if constexpr (SENS) {
                // the same RU on every sdiff (bdf.rs:539-541); the reference's shared ping-pong buffer only moves columns
                // above the order around, which are rewritten before anything reads them
#pragma unroll 1
                for (int qs = 0; qs < NP; ++qs) {
#pragma unroll
                    for (int l = 1; l <= DSB_MAX_ORDER; ++l) {
                        rrow[l] = 1.0;
#pragma unroll
                        for (int s = 0; s < N; ++s) nd[l][s] = -0.0;
                    }
#pragma unroll 1
                    for (int i = 1; i <= k; ++i) {
                        const double i_t = (double)i;
                        double di[N];                   // (loaded before the divisions: a global-slot load hides behind them)
#pragma unroll
                        for (int s = 0; s < N; ++s) di[s] = SDF(qs, i, s);
#pragma unroll
                        for (int l = 1; l <= DSB_MAX_ORDER; ++l) rrow[l] = DSB_DIV(rrow[l] * (i_t - 1.0 - factor * (double)l), i_t);
#pragma unroll
                        for (int j = 1; j <= DSB_MAX_ORDER; ++j) {
                            double ru_ij = rrow[1] * u[j * 6 + 1];
#pragma unroll
                            for (int l = 2; l <= j; ++l) ru_ij = rrow[l] * u[j * 6 + l] + ru_ij;
#pragma unroll
                            for (int s = 0; s < N; ++s) nd[j][s] = di[s] * ru_ij + nd[j][s];
                        }
                    }
#pragma unroll
                    for (int j = 1; j <= DSB_MAX_ORDER; ++j) {
                        if (j <= k) {
#pragma unroll
                            for (int s = 0; s < N; ++s) SDF(qs, j, s) = nd[j][s];
                        }
                    }
                }
                c_sens = new_h * pa.tab.alpha[k];
            }
            c = new_h * pa.tab.alpha[k];
            h = new_h;
            conv.eta = pa.tab.eta_reset_timestep;
            if (!rs_ignore_small && dsb_abs(h) < pa.opt.min_timestep) finish(DSB_STATUS_STEP_SIZE_TOO_SMALL);
            else state = after_rescale;
        }

        // ================= JAC: _jacobian_updates(c, kind) / Bdf::_new's reset_jacobian ===================
        if (run_setup && state == L_JAC) {
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
                LaneLU<N, DsbDivShared> lu;
                double pl[NP > 0 ? NP : 1];
#pragma unroll
                for (int j = 0; j < NP; ++j) pl[j] = SP(j);
                if (jacobian_is_stale) {
                    // df/dy at (state.y, state.t) (quirk Q6), assembled straight into shared memory
                    double yl[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) yl[i] = SY(i);
                    lane_jacobian_to<M>(pa, yl, pl, t, st, [&](int j, int i, double val) { SJ(j, i) = val; });
                    if (M::HAS_MASS) {
                        lane_mass_matrix<M>(pl, t, lu.a);
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int i = 0; i < N; ++i) SMM(j, i) = lu.a[j][i];
                    }
                    jacobian_is_stale = false;
                }
                const double mc = -c;
#pragma unroll
                for (int j = 0; j < N; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        // identity mass when the model has none (op/bdf.rs:141-143)
                        const double m_ji = M::HAS_MASS ? SMM(j, i) : ((i == j) ? 1.0 : 0.0);
                        lu.a[j][i] = SJ(j, i) * mc + m_ji;
                    }
                lu.factor();
                piv_packed = 0;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    piv_packed |= (unsigned long long)lu.piv[j] << (4 * j);
#pragma unroll
                    for (int i = 0; i < N; ++i) SLU(j, i) = lu.a[j][i];
#ifdef DSB_LANE_LU_RCP
                    SM(Lay::O_RCP + j) = lu.rinv[j];
#endif
                }
            }
            state = after_jac;
        }

        // ================= TSTOP: set_stop_time (first) / handle_tstop after an accepted step =============
        if (__any_sync(0xffffffffu, state == L_TSTOP) && state == L_TSTOP) {
            if constexpr (RAG) {                    // solve(final_time): the initial column, before the stop time is set (method.rs:900-901)
                if (first) {
                    double y0c[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) y0c[i] = SY(i);
                    write_column(0, t, y0c);
                    col = 1;
                }
            }
            bool stopped_on_root = false;
            bool reset_now = false;             // a reset was applied at a root: set_stop_time again, then L_REINIT
            if constexpr (NR > 0) {
                // check for a root within the accepted step (bdf.rs:1566-1579), after the step-size update and before
                // the stop time is handled; RootFinder::check_root (root.rs:60-160) with Vector::root_finding
                // (diffsol-la/src/vector/nalgebra_serial.rs:484-504)
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
                        // (bdf.rs:1228-1262), then -- without a reset function -- the state at the root in the next column
                        // (method.rs:493-503) and the end of the solve
                        double yo[N];
                        while (!RAG && !free_running && col < nt && bb.t_eval[col] <= t_root) {
                            interpolate(bb.t_eval[col], yo);
                            write_column(col, bb.t_eval[col], yo);
                            ++col;
                        }
                        interpolate(t_root, yo);
                        double dyg[N];                      // a DAE's reset: state_mut_back also interpolates dy (bdf.rs:1245-1252, 788-810)
                        if constexpr (dsb_model_has_reset<M>::value && M::HAS_MASS) {
                            double pi = 1.0, d_pi = 0.0;
#pragma unroll
                            for (int i = 0; i < N; ++i) dyg[i] = 0.0;
#pragma unroll 1
                            for (int j = 0; j < order; ++j) {
                                const double j_t = (double)j;
                                const double denom = h * (1.0 + j_t);
                                const double w = DSB_DIV(t_root - (t - h * j_t), denom);
                                const double dw = DSB_DIV(1.0, denom);
                                const double new_d_pi = d_pi * w + pi * dw;
                                pi *= w;
                                d_pi = new_d_pi;
#pragma unroll
                                for (int i = 0; i < N; ++i) dyg[i] = d_pi * SD(j + 1, i) + dyg[i];
                            }
                        }
                        if (!free_running) t = t_root;      // state_mut_back; the harness loop leaves the state at the end of the step
                        bool ended = true;
                        if constexpr (dsb_model_has_reset<M>::value) {
                            if (!free_running) {
                                // has_reset (method.rs:783-797): apply_reset (state.rs:246-270: y <- reset(y, t),
                                // dy <- f(y, t)), then a new stop time and on with the integration -- or TstopReached
                                double yr[N], dyr[N];
                                M::reset(yo, pl, t, yr);
                                int reset_status = DSB_STATUS_OK;
                                if constexpr (M::HAS_MASS) {
                                    // state.apply_reset_with_mass (state.rs:279-306): set_consistent with a Newton solver WITHOUT
                                    // line search, from the reset y and the dy interpolated at the root; InitOp and the mass
                                    // matrix are evaluated at problem.t0 (state.rs:114-119)
                                    LaneConvergence ic_conv;
                                    ic_conv.tol = pa.opt.nonlinear_solver_tolerance;
                                    ic_conv.eta = pa.tab.eta_reset;
                                    ic_conv.max_iter = pa.opt.ic_max_newton_iterations;
                                    ic_conv.reset();
#pragma unroll
                                    for (int i = 0; i < N; ++i) dyr[i] = dyg[i];
                                    reset_status = lane_consistent_solve<M>(pa, pl, yr, dyr,
                                        [&](const double (&x)[N], double (&out)[N]) { M::rhs(x, pl, pa.t0, out); st.v[DSB_STAT_RHS_CALLS] += 1; },
                                        [&](double (&J)[N][N]) {
                                            lane_jacobian_to<M>(pa, yr, pl, pa.t0, st, [&](int j, int i, double val) { J[j][i] = val; });
                                        }, ic_conv, true, false);
                                } else {
                                    M::rhs(yr, pl, t, dyr);
                                    st.v[DSB_STAT_RHS_CALLS] += 1;
                                }
#pragma unroll
                                for (int i = 0; i < N; ++i) { SY(i) = yr[i]; SYP(i) = dyr[i]; }
                                root_found = -1;
                                if (reset_status != DSB_STATUS_OK) finish(reset_status);
                                else if (t < bb.t_eval[nt - 1]) { reset_now = true; stopped_on_root = false; }
                                else finish(DSB_STATUS_OK);                        // TstopReached
                                ended = false;
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
            int next = first ? L_PREDICT : L_OUTPUT;
            int r = 0;
            bool check = has_tstop && !stopped_on_root;
            if (reset_now) { next = L_REINIT; check = true; has_tstop = true; tstop = bb.t_eval[nt - 1]; }
            if (first) {
                check = !free_running;
                if (free_running) next = L_OUTPUT;
                else { has_tstop = true; tstop = bb.t_eval[nt - 1]; }
            }
            if (check) {                                 // one call site for handle_tstop (code size)
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

        // ================= OUTPUT: dense output at every t_eval passed (method.rs:761-764, 822-848) =======
        if (__any_sync(0xffffffffu, state == L_OUTPUT) && state == L_OUTPUT) {
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
                // interpolate (bdf.rs:767-782, 1080-1106)
                const bool is_forward = h > 0.0;
                if ((is_forward && tq > t) || (!is_forward && tq < t)) { status = DSB_STATUS_INTERPOLATION_TIME_AFTER_CURRENT; break; }
                double yo[N];
                interpolate(tq, yo);
                write_column(col, tq, yo);
                if constexpr (SENS) {                   // interpolate_sens (bdf.rs:1162-1215): interpolate_from_diff on every sdiff
#pragma unroll 1
                    for (int qs = 0; qs < NP; ++qs) {
                        double time_factor = 1.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) yo[i] = SDF(qs, 0, i);
#pragma unroll 1
                        for (int j = 0; j < order; ++j) {
                            const double j_t = (double)j;
                            time_factor *= DSB_DIV(tq - (t - h * j_t), h * (1.0 + j_t));
#pragma unroll
                            for (int i = 0; i < N; ++i) yo[i] = time_factor * SDF(qs, j + 1, i) + yo[i];
                        }
#pragma unroll
                        for (int i = 0; i < N; ++i) bb.ss[(((int64_t)col * NP + qs) * N + i) * B + inst] = yo[i];
                    }
                }
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

        // ================= PREDICT: _predict_forward + start of a Newton solve ============================
        DSB_PROF_BLOCK(5, state == L_PREDICT)
        if (__any_sync(0xffffffffu, state == L_PREDICT) && state == L_PREDICT) {
            if (repredict || SENS) {        // (sensitivities: psi_neg_y0 was reused by the sensitivity solves; the same values again)
                double yp[N];
#pragma unroll
                for (int i = 0; i < N; ++i) yp[i] = 0.0;
#pragma unroll 1
                for (int j = 0; j <= order; ++j) {
#pragma unroll
                    for (int i = 0; i < N; ++i) yp[i] += SD(j, i);
                }
#pragma unroll
                for (int i = 0; i < N; ++i) psi_neg_y0[i] = pa.tab.gamma[1] * SD(1, i);
#pragma unroll 1
                for (int j = 2; j <= order; ++j) {
                    const double g = pa.tab.gamma[j];
#pragma unroll
                    for (int i = 0; i < N; ++i) psi_neg_y0[i] = g * SD(j, i) + psi_neg_y0[i];
                }
                const double a = pa.tab.alpha[order];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    psi_neg_y0[i] *= a;
                    psi_neg_y0[i] -= yp[i];
                    SYP(i) = yp[i];
                }
                t_predict = t + h;
                sens_predict_all();
            }
            state = L_NEWTON;
            if (pending_etf) {
                pending_etf = false;
                st.v[DSB_STAT_ERROR_TEST_FAILURES] += 1;
                if (st.v[DSB_STAT_ERROR_TEST_FAILURES] - old_num_error_test_failures >= pa.opt.max_error_test_failures)
                    finish(DSB_STATUS_TOO_MANY_ERROR_TEST_FAILURES);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) y_cur[i] = SYP(i);
            conv.reset();
        }

        // ================= NEWTON: one iteration (newton.rs:13-36, line_search.rs:48-69) ==================
        // A lane that needs another iteration takes it in the same trip (up to newton_passes of them): the trip's other blocks
        // -- the per-step chain is the larger part of the loop body -- then run once per STEP of most lanes, not once per
        // iteration, with correspondingly more lanes each time.  The loop is rolled: one copy of the block in the code.
#pragma unroll 1
        for (int pass = 0; pass < newton_passes; ++pass) {
        DSB_PROF_BLOCK(7, state == L_NEWTON)
        if (!__any_sync(0xffffffffu, state == L_NEWTON)) break;
        if (state == L_NEWTON) {
            double pl[NP > 0 ? NP : 1];
#pragma unroll
            for (int j = 0; j < NP; ++j) pl[j] = SP(j);
            double delta[N];
            bool sens_eq = false;
            if constexpr (SENS) sens_eq = eq > 0;
            if (sens_eq) {
                if constexpr (SENS) {
                    // BdfCallable<SensEquations>::call_inplace (op/bdf.rs:240-256) on SensRhs::call_inplace
                    // (sens_equations.rs:168-174): J(y_predict) s + f_p e_q, with the sensitivity residual's own c
                    double ypl_[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) ypl_[i] = SYP(i);
                    M::jac_mul(ypl_, pl, t_predict, y_cur, delta);
                    st.v[DSB_STAT_RHS_JAC_MULS] += 1;
                    const double mc = -c_sens;
                    double tmp[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        delta[i] += SFP(eq - 1, i);
                        tmp[i] = y_cur[i] + psi_neg_y0[i];
                    }
                    if (M::HAS_MASS) {
                        M::mass(tmp, pl, t_predict, mc, delta);
                    } else {
#pragma unroll
                        for (int i = 0; i < N; ++i) delta[i] = tmp[i] + mc * delta[i];
                    }
                }
            } else {
            M::rhs(y_cur, pl, t_predict, delta);
            st.v[DSB_STAT_RHS_CALLS] += 1;
            {
                double tmp[N];
#pragma unroll
                for (int i = 0; i < N; ++i) tmp[i] = y_cur[i] + psi_neg_y0[i];
                const double mc = -c;
                if (M::HAS_MASS) {
                    M::mass(tmp, pl, t_predict, mc, delta);
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) delta[i] = tmp[i] + mc * delta[i];
                }
            }
            }
            LaneLU<N, DSB_NEWTON_DIV> lu;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                lu.piv[j] = (int)((piv_packed >> (4 * j)) & 15ull);
#pragma unroll
                for (int i = 0; i < N; ++i) lu.a[j][i] = SLU(j, i);
#ifdef DSB_LANE_LU_RCP
                lu.rinv[j] = SM(Lay::O_RCP + j);
#endif
            }
#ifdef DSB_LANE_LU_RCP
            if (!lu.template solve<true>(delta)) {
#else
            if (!lu.solve(delta)) {
#endif
                newton_ok = false; state = L_POST;              // LuSolveFailed
            } else {
                double ypl[N];
#pragma unroll
                for (int i = 0; i < N; ++i) { y_cur[i] -= delta[i]; ypl[i] = sens_eq ? SPR(eq - 1, i) : SYP(i); }
                // Newton norm weights use the PREDICTOR (line_search.rs:67, convergence.rs:64-66)
                const double acc = weighted_sum(delta, ypl);
                const double norm = dsb_sqrt(DSB_DIV_N(acc));
                // Convergence::check_new_iteration (convergence.rs:68-139) with its pow() hoisted to one call site
                conv.niter += 1;
                const bool have_rate = conv.has_old_norm;
                double px, py;
#ifndef DSB_NO_HOST_TABLES     // A/B switch: 61.5 -> 59.4 ms per 10^6 Robertson instances (profiles/r2_lane_kernel_ab.log)
                if (have_rate) { px = DSB_DIV(norm, conv.old_norm); py = conv.niter <= 32 ? pa.tab.inv_int[(conv.niter - 1) & 31] : DSB_DIV(1.0, (double)(conv.niter - 1)); }
#else
                if (have_rate) { px = DSB_DIV(norm, conv.old_norm); py = DSB_DIV(1.0, (double)(conv.niter - 1)); }
#endif
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
        // ================= POST: a Newton solve ended (bdf.rs:1338-1563) ==================================
        DSB_PROF_BLOCK(9, state == L_POST)
        // Sensitivities: a lane passes through NEWTON (1 + np) x (iterations) times per step and through the END of POST (error
        // test, _update_diff on every difference array, then TSTOP / OUTPUT / PREDICT) once; left alone those blocks run on
        // every trip with ~3 lanes (ncu, profiles/r2_sens_kernel_ncu_summary.txt).  A lane whose last sensitivity solve has
        // converged therefore WAITS in front of POST until a quorum (or half of the active lanes) is there too -- waiting
        // changes when a lane runs, never its arithmetic; with two waiting groups (this one and the slow pool) that each
        // release at half of the active lanes, all lanes can never wait at once.
        bool hold_post = false;
        if constexpr (SENS) {
            const bool at_end = state == L_POST && newton_ok && eq == NP;
            const int n_end = __popc(__ballot_sync(0xffffffffu, at_end));
            hold_post = at_end && !(n_end >= quorum || 2 * n_end >= n_active);
        }
        if (__any_sync(0xffffffffu, state == L_POST && !hold_post) && state == L_POST && !hold_post) {
            bool run_main = true;
            if constexpr (SENS) {
                // sensitivity_solve (bdf.rs:934-989) after a successful main solve: one Newton solve per parameter on the
                // same LU and the same Convergence; a failed sensitivity solve is a failed step whose iterations are NOT counted
                // (the `?` returns before the statistics line)
                if (eq == 0 || newton_ok) st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
                if (newton_ok) {
                    if (eq == 0) {
                        double ypl_[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) { SYC(i) = y_cur[i]; ypl_[i] = SYP(i); }
                        sens_update_state(ypl_, t_predict);
                    } else {
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            SSS(eq - 1, i) = y_cur[i];
                            double dl = y_cur[i];
                            dl -= SPR(eq - 1, i);
                            SDL(eq - 1, i) = dl;
                        }
                    }
                    if (eq < NP) {
                        sens_setup(eq);
                        eq += 1;
                        state = L_NEWTON;
                        run_main = false;
                    } else {
                        eq = 0;
#pragma unroll
                        for (int i = 0; i < N; ++i) y_cur[i] = SYC(i);
                    }
                } else {
                    eq = 0;
                }
            } else {
                st.v[DSB_STAT_NONLINEAR_SOLVER_ITERATIONS] += conv.niter;
            }
            if (!run_main) {
                // on to the next sensitivity solve
            } else if (newton_ok) {
                const int ord = order;
                double d[N];
#pragma unroll
                for (int i = 0; i < N; ++i) d[i] = y_cur[i] - SYP(i);
                {   // error_control: ||d||^2_w(state.y) * error_const2[order - 1]
                    double yl[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) yl[i] = SY(i);
                    const double acc = weighted_sum(d, yl);
                    const double err = DSB_DIV_N(acc) * pa.tab.error_const2[ord - 1];
                    error_norm = (0.0 < err) ? err : 0.0;
                }
                if constexpr (SENS) {               // NB error_const2[order] for the sensitivities (bdf.rs:844-858)
                    if (pa.sens_error_control) {
#pragma unroll 1
                        for (int qs = 0; qs < NP; ++qs) {
                            double x[N], ref[N];
#pragma unroll
                            for (int i = 0; i < N; ++i) { x[i] = SDL(qs, i); ref[i] = SSS(qs, i); }
                            const double es = sens_weighted_norm(x, ref) * pa.tab.error_const2[ord];
                            error_norm = (error_norm < es) ? es : error_norm;
                        }
                    }
                }
#ifndef DSB_NO_HOST_TABLES     // A/B switch: 61.5 -> 59.4 ms per 10^6 Robertson instances (profiles/r2_lane_kernel_ab.log)
                if (conv.niter < 32) {
                    safety = pa.tab.safety[conv.niter];             // 0.9 (2 m + 1) / (2 m + niter), divided on the host
                } else
#endif
                {
                    const double maxiter = (double)conv.max_iter;
                    const double niter = (double)conv.niter;
                    safety = DSB_DIV(0.9 * (2.0 * maxiter + 1.0), 2.0 * maxiter + niter);
                }
                if (error_norm <= 1.0) {
                    // ---- accepted: _update_diff, state.y <- PREDICTOR (quirk Q1) ----
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        SD(ord + 2, i) = d[i] - SD(ord + 1, i);
                        SD(ord + 1, i) = d[i];
                    }
#pragma unroll 1
                    for (int j = ord; j >= 0; --j) {
#pragma unroll
                        for (int i = 0; i < N; ++i) SD(j, i) = SD(j, i) + 1.0 * SD(j + 1, i);
                    }
                    if constexpr (Lay::SDIFF_GLOBAL) {      // the same update with one array's columns loaded together (global slot)
#pragma unroll 1
                        for (int qs = 0; qs < NP; ++qs) {
                            double col[DSB_MAX_ORDER + 1][N], o1[N], prev[N];
#pragma unroll
                            for (int j = 0; j <= DSB_MAX_ORDER; ++j)
                                if (j <= ord) {
#pragma unroll
                                    for (int i = 0; i < N; ++i) col[j][i] = SDF(qs, j, i);
                                }
#pragma unroll
                            for (int i = 0; i < N; ++i) o1[i] = SDF(qs, ord + 1, i);
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double dl = SDL(qs, i);
                                SDF(qs, ord + 2, i) = dl - o1[i];
                                SDF(qs, ord + 1, i) = dl;
                                prev[i] = dl;
                            }
#pragma unroll
                            for (int j = DSB_MAX_ORDER; j >= 0; --j)
                                if (j <= ord) {
#pragma unroll
                                    for (int i = 0; i < N; ++i) { prev[i] = col[j][i] + 1.0 * prev[i]; SDF(qs, j, i) = prev[i]; }
                                }
                        }
                    } else if constexpr (SENS) {           // _update_diff on every sdiff (bdf.rs:629-633)
#pragma unroll 1
                        for (int qs = 0; qs < NP; ++qs) {
#pragma unroll
                            double prev[N];
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double dl = SDL(qs, i);
                                SDF(qs, ord + 2, i) = dl - SDF(qs, ord + 1, i);
                                SDF(qs, ord + 1, i) = dl;
                                prev[i] = dl;
                            }
#pragma unroll 1
                            for (int j = ord; j >= 0; --j) {
#pragma unroll
                                for (int i = 0; i < N; ++i) { prev[i] = SDF(qs, j, i) + 1.0 * prev[i]; SDF(qs, j, i) = prev[i]; }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) SY(i) = SYP(i);
                    t = t_predict;
                    st.v[DSB_STAT_STEPS] += 1;
                    ju.step();
                    has_prev_error = true; prev_error_norm = error_norm;
                    n_equal_steps += 1;
                    accepted = true;
                    state = (n_equal_steps > ord) ? L_SELECT : L_TSTOP;
                } else {
                    accepted = false;               // error test failed: the new step size needs a pow(), see SELECT
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
#ifdef DSB_LANE_PROFILE
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) atomicAdd(work_counter + 1 + k, prof[k]);
    }
#endif
#undef DSB_PROF_BLOCK
#undef SM
#undef DSB_DIV
#undef DSB_DIV_N
#undef SD
#undef SJ
#undef SMM
#undef SLU
#undef SY
#undef SYP
#undef SP
#undef SDF
#undef SSS
#undef SDL
#undef SFP
#undef SPR
#undef SPS
#undef SYC
}
