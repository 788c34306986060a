// dsb_band_init_kernel.cuh -- `OdeSolverState::new_and_consistent` for banded DAEs of medium size (n > 16, singular
// mass matrix inside the declared band): what dsb_init_kernel.cuh does for the on-chip lane kernels, with the
// instance's vectors and matrices in the lane's global-memory column (dsb_band_bdf_kernel.cuh) and the InitOp
// Jacobian factored by the per-lane band LU (dsb_band_lu.cuh).
//
// Restates (paths relative to /root/reference/crates/diffsol/src):
//   new_without_initialise   ode_solver/state.rs:1086-1124    y = init(p, t0); dy = f(y, t0)
//   set_consistent           ode_solver/state.rs:84-162 + op/init.rs:14-131
//                            Newton with BacktrackingLineSearch (diffsol-nl/src/line_search.rs:115-201)
//                            on F(du, v) = -M_u du + f(u, v); g(u, v)
// The InitOp Jacobian (-M_u | f_v ; 0 | g_v) and neg_mass (-M_u | 0 ; 0 | 0) have the band of M and df/dy; only
// products with exactly zero entries are skipped, so the results are those of the dense restatement.
// One lane per instance (grid-stride); y, dy, the counters and the status go to the batch-major arrays the
// integrator kernel starts from.  The initial step size is computed by the integrator's FETCH block.
#pragma once
#include "dsb_band_bdf_kernel.cuh"

template <class M, int T>
__global__ void __launch_bounds__(T) dsb_band_init_kernel(const __grid_constant__ DsbProblemArgs pa,
                                                          const __grid_constant__ DsbBatchBuffers bb,
                                                          const __grid_constant__ DsbBandMeta meta,
                                                          double* __restrict__ ws) {
    typedef BandBdfLayout<M, T> Lay;
    constexpr int U2 = BandUnroll<T>::U2, U4 = BandUnroll<T>::U4;
    typedef LaneBandLU<M::N, Lay::KL, Lay::KU, DsbDivInline, U2> BLU;
    constexpr int N = Lay::N, NP = Lay::NP, KL = Lay::KL, KU = Lay::KU, KV = Lay::KV, LDJ = Lay::LDJ, LDAB = Lay::LDAB;
    // the integrator's words, reused: y, dy (D[1]), x = (du, v) iterate, yerr, delta, x0, delta0, InitOp.y0
    constexpr int O_YV = Lay::O_Y, O_DY = Lay::O_D + N, O_X = Lay::O_YC, O_YERR = Lay::O_YP, O_DELTA = Lay::O_DL,
                  O_X0 = Lay::O_PSI, O_DELTA0 = Lay::O_D + 2 * N, O_Y0W = Lay::O_D + 3 * N;
    const size_t LS = (size_t)gridDim.x * blockDim.x;
    double* const g = ws + ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
#define G(w) g[(size_t)(w) * LS]
#define GJ(j, r) G(Lay::O_J + (j) * LDJ + (r))
#define GM(j, r) G(Lay::O_M + (j) * LDJ + (r))
#define GAB(j, r) G(Lay::O_LU + (j) * LDAB + (r))
    const BandVec vY{g + (size_t)O_YV * LS, LS}, vY0W{g + (size_t)O_Y0W * LS, LS};
    const int64_t B = pa.nbatch;
    for (int64_t inst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; inst < B; inst += (int64_t)LS) {
        double pl[NP > 0 ? NP : 1];
#pragma unroll
        for (int j = 0; j < NP; ++j) pl[j] = bb.params[(int64_t)j * B + inst];
        LaneStats st;
        st.clear();
        int status = DSB_STATUS_OK;
        const double t0 = pa.t0;
        // ||x||^2_w(ref) over words of the lane's column (vector/nalgebra_serial.rs:395-408)
        auto weighted_norm = [&](int ox, int oref) -> double {
            double acc = 0.0;
#pragma unroll U4
            for (int i = 0; i < N; ++i) {
                const double term = G(ox + i) / (dsb_abs(G(oref + i)) * pa.rtol + meta.atol[i]);
                acc += term * term;
            }
            return acc / (double)N;
        };
#pragma unroll U2
        for (int i = 0; i < N; ++i) G(O_YV + i) = M::init_i(i, pl, t0);
#pragma unroll U2
        for (int i = 0; i < N; ++i) G(O_DY + i) = M::rhs_i(i, vY, pl, t0);
        st.v[DSB_STAT_RHS_CALLS] += 1;

        // mass matrix at t0, column j = M e_j with beta = 0 (op/linear_op.rs:42-51); algebraic <=> zero diagonal
        int nalg = 0;
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const BandUnitVec ej{j};
#pragma unroll
            for (int r = 0; r < LDJ; ++r) {
                const int i = j + r - KU;
                const double m = (i >= 0 && i < N) ? M::mass_i(i, ej, pl, t0, 0.0, 0.0) : 0.0;
                GM(j, r) = m;
                if (i == j && m == 0.0) nalg += 1;
            }
        }
        if (nalg > 0) {
            auto is_alg = [&](int i) -> bool { return GM(i, KU) == 0.0; };
            // df/dy at (y0, t0) (InitOp::new, op/init.rs:22-76): the integrator's coloured assembly
            st.v[DSB_STAT_RHS_MATRIX_EVALS] += 1;
            for (int e = 0; e < LDJ * N; ++e) G(Lay::O_J + e) = 0.0;
            const bool one_colour_per_column = pa.ncolors == N;
#pragma unroll 1
            for (int cc = 0; cc < pa.ncolors; ++cc) {
                const BandColourSeed seed{meta.colmeta, cc};
                st.v[DSB_STAT_RHS_JAC_MULS] += 1;
                const int i0 = one_colour_per_column ? (cc - KU < 0 ? 0 : cc - KU) : 0;
                const int i1 = one_colour_per_column ? (cc + KL > N - 1 ? N - 1 : cc + KL) : N - 1;
#pragma unroll 1
                for (int i = i0; i <= i1; ++i) {
                    const double val = M::jac_mul_i(i, vY, pl, t0, seed);
#pragma unroll
                    for (int d = -KL; d <= KU; ++d) {               // column j = i + d
                        const int j = i + d;
                        if (j >= 0 && j < N) {
                            const int32_t m = meta.colmeta[j];
                            if ((m & 0xffff) == cc && ((m >> (16 + KU - d)) & 1)) GJ(j, KU - d) = val;
                        }
                    }
                }
            }
            // entry (i, j) of neg_mass: -M_u on the differential block, zero elsewhere
            auto neg_mass = [&](int i, int j, bool alg_i) -> double {
                return (!alg_i && !is_alg(j)) ? GM(j, KU + i - j) * -1.0 : 0.0;
            };
            // fun(x) -> delta: y0[alg] = x[alg]; out = f(y0, t0); out = neg_mass x + out (column sweep: ascending j)
            auto fun = [&]() {
#pragma unroll U2
                for (int i = 0; i < N; ++i) if (is_alg(i)) G(O_Y0W + i) = G(O_X + i);
#pragma unroll 1
                for (int i = 0; i < N; ++i) {
                    double o = M::rhs_i(i, vY0W, pl, t0);
                    const bool alg_i = is_alg(i);
#pragma unroll
                    for (int d = -KL; d <= KU; ++d) {
                        const int j = i + d;
                        if (j >= 0 && j < N) o = neg_mass(i, j, alg_i) * G(O_X + j) + o;
                    }
                    G(O_DELTA + i) = o;
                }
                st.v[DSB_STAT_RHS_CALLS] += 1;
            };
#pragma unroll U2
            for (int i = 0; i < N; ++i) {
                const double v = is_alg(i) ? G(O_YV + i) : G(O_DY + i);
                G(O_X + i) = v; G(O_YERR + i) = v; G(O_Y0W + i) = G(O_YV + i);
            }
            LaneConvergence conv;
            conv.tol = pa.opt.nonlinear_solver_tolerance;
            conv.eta = pa.tab.eta_reset;
            conv.max_iter = pa.opt.ic_max_newton_iterations;
            conv.old_norm = 0.0;
            conv.reset();
            const double tau = pa.opt.ic_step_reduction_factor, c_armijo = pa.opt.ic_armijo_constant;
            const double steptol = pa.tab.ic_steptol;
            const int ls_max_iter = pa.opt.ic_max_linesearch_iterations;
            bool ok = false;
            for (int setup = 0; setup < pa.opt.ic_max_linear_solver_setups && status == DSB_STATUS_OK && !ok; ++setup) {
                // jac = (-M_u | f_v ; 0 | g_v) in band storage, then its LU
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const bool alg_j = is_alg(j);
#pragma unroll
                    for (int r = 0; r < LDAB; ++r) {
                        const int i = j + r - KV;
                        double v = 0.0;
                        if (r >= KL && i >= 0 && i < N) v = alg_j ? GJ(j, r - KL) : neg_mass(i, j, is_alg(i));
                        GAB(j, r) = v;
                    }
                }
                BLU::factor(g, LS, Lay::O_LU, Lay::O_PIV);
                conv.reset();
                for (int i = 0; i < N; ++i) G(O_DELTA + i) = 0.0;
                double ls_norm = 1.0;
                int result = -1;   // 0 ok, 1 max iterations, 2 other error
                for (int it = 0; it < conv.max_iter && result < 0; ++it) {
                    int res = LANE_CONTINUE;
                    bool have_res = false;
                    if (pa.opt.ic_use_linesearch) {
                        if (conv.niter == 0) {
                            fun();
                            if (!BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, O_DELTA)) { result = 2; break; }
                            ls_norm = dsb_sqrt(weighted_norm(O_DELTA, O_YERR));
                            if (conv.check_norm(ls_norm) == LANE_CONVERGED) {
#pragma unroll U4
                                for (int i = 0; i < N; ++i) G(O_X + i) -= G(O_DELTA + i);
                                res = LANE_CONVERGED; have_res = true;
                            }
                        }
                        if (!have_res) {
#pragma unroll U4
                            for (int i = 0; i < N; ++i) { G(O_X0 + i) = G(O_X + i); G(O_DELTA0 + i) = G(O_DELTA + i); }
                            const double norm = ls_norm;
                            const double phi0 = norm * norm * 0.5, two_phi0 = norm * norm;
                            const double min_alpha = steptol / norm;
                            double alpha = 1.0;
                            int ls_status = 1;
                            for (int li = 0; li < ls_max_iter; ++li) {
#pragma unroll U4
                                for (int q = 0; q < N; ++q) G(O_X + q) = (-alpha) * G(O_DELTA0 + q) + G(O_X + q);
                                fun();
                                if (!BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, O_DELTA)) { ls_status = 2; break; }
                                const double new_norm = dsb_sqrt(weighted_norm(O_DELTA, O_YERR));
                                const double phi1 = new_norm * new_norm * 0.5;
                                if (phi1 <= phi0 - c_armijo * alpha * two_phi0) {
                                    ls_norm = new_norm;
                                    res = conv.check_norm(new_norm); have_res = true; ls_status = 0;
                                    break;
                                }
                                if (alpha < min_alpha) { ls_status = 2; break; }
                                alpha *= tau;
#pragma unroll U4
                                for (int q = 0; q < N; ++q) G(O_X + q) = G(O_X0 + q);
                            }
                            if (ls_status != 0) { result = 2; break; }
                        }
                    } else {
                        fun();
                        if (!BLU::solve(g, LS, Lay::O_LU, Lay::O_PIV, O_DELTA)) { result = 2; break; }
#pragma unroll U4
                        for (int i = 0; i < N; ++i) G(O_X + i) -= G(O_DELTA + i);
                        res = conv.check_new_iteration(dsb_sqrt(weighted_norm(O_DELTA, O_YERR)));
                    }
                    if (res == LANE_CONVERGED) result = 0;
                    else if (res == LANE_DIVERGED) result = 2;
                }
                if (result < 0) result = 1;
                if (result == 0) ok = true;
                else if (result == 2) status = DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
                else {
#pragma unroll U4
                    for (int i = 0; i < N; ++i) G(O_YERR + i) = G(O_X + i);
                }
            }
            if (!ok) status = DSB_STATUS_INITIAL_CONDITION_DID_NOT_CONVERGE;
            if (status == DSB_STATUS_OK) {
#pragma unroll U2
                for (int i = 0; i < N; ++i) {
                    if (is_alg(i)) { G(O_YV + i) = G(O_X + i); G(O_DY + i) = 0.0; }
                    else G(O_DY + i) = G(O_X + i);
                }
            }
        }
#pragma unroll U2
        for (int i = 0; i < N; ++i) {
            bb.y0[(int64_t)i * B + inst] = G(O_YV + i);
            bb.dy0[(int64_t)i * B + inst] = G(O_DY + i);
        }
        bb.status[inst] = status;
#pragma unroll
        for (int s = 0; s < DSB_NSTATS; ++s) bb.stats[(int64_t)s * B + inst] = st.v[s];
    }
#undef G
#undef GJ
#undef GM
#undef GAB
}
