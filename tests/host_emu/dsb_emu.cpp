// dsb_emu.cpp -- TEST INFRASTRUCTURE: single-lane host build of the product's lane kernels (see cuda_shim.h).
// `emu_solve` runs `problem.<method>().solve_dense(t_eval)` (or the free-running step/interpolate loop) for each
// instance through the SAME kernel source the sm_100a library is built from, one instance per "launch", and hands
// back what the C ABI's *_host entry points would.  Used by tests/test_host_emu_*.py to compare kernel logic with the
// oracle without a GPU; never part of the product.
#include "cuda_shim.h"

#include <cstdlib>
#include <vector>

#include "../../diffsol_b200/csrc/dsb_host_setup.h"
#include "../../diffsol_b200/csrc/dsb_band_bdf_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_band_init_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_band_sdirk_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_bdf_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_init_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_sdirk_kernel.cuh"
#include "../../diffsol_b200/csrc/dsb_wband_bdf_kernel.cuh"

double dsb_lane_smem[1 << 20];      // the "shared memory" of the emulated block (word w of lane 0 at w * THREADS)

namespace {

struct EmuDims { dsb_problem* p; template <class M> void operator()() { p->n = M::N; p->np = M::NP; p->has_mass = M::HAS_MASS; } };

struct EmuCall {
    const dsb_problem* pr; int method, kernel, free_running;
    const double* params; int64_t B; const double* t_eval; int nt;
    double* ys; int64_t* stats; int32_t* status; double* fin; int32_t* root_idx; int32_t* ncols;
    int rc;

    template <class M> void operator()() {
        constexpr int N = M::N, NP = M::NP;
        DsbProblemArgs pa;
        int probes = 0;
        std::vector<int32_t> color_full; std::vector<uint8_t> nz_full;
        if (dsb_host::fill_problem_args(*pr, 1, nt, &pa, &probes, &color_full, &nz_full) != DSB_OK) { rc = DSB_BAD_ARG; return; }
        pa.free_running = free_running;
        dsb_host::build_tableau(method, &pa.rk);
        pa.quorum = DSB_DEFAULT_QUORUM; pa.newton_passes = DSB_DEFAULT_NEWTON_PASSES;
        if (const char* q = getenv("DSB_WBAND_FORCE_REDO")) pa.reserved1 = atoi(q);
        std::vector<double> atol_full(N);
        for (int i = 0; i < N; ++i) atol_full[i] = pr->atol.size() == 1 ? pr->atol[0] : pr->atol[i];
        constexpr int NOUT = dsb_model_nout<M>::value;
        std::vector<double> y0(N), dy0(N), h0(1), ysb((size_t)nt * NOUT), fin_t(1), fin_h(1);
        std::vector<int32_t> st(DSB_NSTATS), status1(1), fin_order(1), ridx(1), nc(1);
        for (int64_t b = 0; b < B; ++b) {
            DsbBatchBuffers bb{};
            bb.params = params + b * NP; bb.t_eval = t_eval; bb.y0 = y0.data(); bb.dy0 = dy0.data(); bb.h0 = h0.data();
            bb.ys = ysb.data(); bb.stats = st.data(); bb.status = status1.data();
            bb.fin_t = fin_t.data(); bb.fin_h = fin_h.data(); bb.fin_order = fin_order.data();
            bb.root_idx = ridx.data(); bb.ncols = nc.data();
            for (auto& v : ysb) v = dsb_from_bits(~0ull);
            for (auto& v : st) v = 0;
            status1[0] = -1; ridx[0] = -1; nc[0] = nt; fin_t[0] = fin_h[0] = 0.0; fin_order[0] = 0;
            unsigned long long work_counter = 0;
            bool ran = false;
            if (kernel == 1) {
                if constexpr (N <= 16) {
                    if (method == DSB_METHOD_BDF) {
                        blockDim.x = BdfLayout<M>::THREADS;
                        dsb_init_kernel<M>(pa, bb, 1);
                        dsb_bdf_solve_dense_kernel<M>(pa, bb, &work_counter);
                    } else {
                        blockDim.x = SdirkLayout<M>::THREADS;
                        dsb_init_kernel<M>(pa, bb, pa.rk.order);
                        dsb_sdirk_solve_dense_kernel<M>(pa, bb, &work_counter);
                    }
                    ran = true;
                }
            } else if (kernel == 3) {
                if constexpr (dsb_declares_band<M>::value && dsb_is_componentwise<M>::value && N > 16) {
                    constexpr int T = DSB_BAND_THREADS_SMALL;
                    typedef BandSdirkLayout<M, T> Lay;           // WORDS covers both kernels and the initialisation
                    std::vector<int32_t> colmeta;
                    if (!dsb_host::band_column_meta<M>(pa.t0, pa.use_coloring != 0, color_full.empty() ? nullptr : color_full.data(),
                                                       Lay::KL, Lay::KU, &colmeta)) { rc = DSB_ERR; return; }
                    const DsbBandMeta meta{atol_full.data(), colmeta.data()};
                    blockDim.x = Lay::THREADS;
                    std::vector<double> ws((size_t)Lay::WORDS * Lay::THREADS);
                    if constexpr (M::HAS_MASS) dsb_band_init_kernel<M, T>(pa, bb, meta, ws.data());
                    if (method == DSB_METHOD_BDF) dsb_band_bdf_solve_dense_kernel<M, T>(pa, bb, meta, ws.data(), &work_counter);
                    else dsb_band_sdirk_solve_dense_kernel<M, T>(pa, bb, meta, ws.data(), &work_counter);
                    ran = true;
                }
            } else if (kernel == 4) {
                // warp-per-instance banded kernel with ONE lane per warp (DSB_WLANES = 1 on the host): same source, same
                // arithmetic; with one instance the instance-major result block is the batch-major one
                if constexpr (dsb_declares_band<M>::value && dsb_is_componentwise<M>::value && N > 16 && !dsb_model_has_reset<M>::value) {
                    typedef WBandLayout<M> LayW;
                    if (method == DSB_METHOD_BDF) {
                        std::vector<int32_t> colmeta;
                        if (!dsb_host::band_column_meta<M>(pa.t0, pa.use_coloring != 0, color_full.empty() ? nullptr : color_full.data(),
                                                           LayW::KL, LayW::KU, &colmeta)) { rc = DSB_ERR; return; }
                        const DsbBandMeta meta{atol_full.data(), colmeta.data()};
                        if constexpr (M::HAS_MASS) {
                            constexpr int T = DSB_BAND_THREADS_SMALL;
                            typedef BandSdirkLayout<M, T> Lay;
                            blockDim.x = Lay::THREADS;
                            std::vector<double> ws((size_t)Lay::WORDS * Lay::THREADS);
                            dsb_band_init_kernel<M, T>(pa, bb, meta, ws.data());
                        }
                        blockDim.x = LayW::THREADS;
                        std::vector<double> slots((size_t)LayW::G_WORDS * LayW::WARPS + 16);
                        dsb_wband_bdf_solve_dense_kernel<M>(pa, bb, meta, slots.data(), bb.ys, &work_counter);
                        ran = true;
                    }
                }
            }
            if (!ran) { rc = DSB_ERR; return; }
            for (int k = 0; k < nt; ++k)
                for (int i = 0; i < NOUT; ++i) ys[(b * nt + k) * NOUT + i] = ysb[(size_t)k * NOUT + i];
            for (int s = 0; s < DSB_NSTATS; ++s) stats[b * DSB_NSTATS + s] = st[s] + (s == DSB_STAT_RHS_JAC_MULS ? probes : 0);
            status[b] = status1[0];
            fin[b * 3 + 0] = fin_t[0]; fin[b * 3 + 1] = fin_h[0]; fin[b * 3 + 2] = (double)fin_order[0];
            root_idx[b] = ridx[0]; ncols[b] = nc[0];
        }
        rc = DSB_OK;
    }
};

// forward sensitivities: the on-chip BDF lane kernel instantiated for DsbWithSens<M> (dsb_inst.cu: SensLauncher)
struct EmuSensCall {
    const dsb_problem* pr; int method; int free_running;
    const double* params; int64_t B; const double* t_eval; int nt;
    double* ys; double* sens; int64_t* stats; int32_t* status;
    int rc;
    template <class M> void operator()() {
        constexpr int N = M::N, NP = M::NP;
        if constexpr (N <= 16 && dsb_model_has_sens<M>::value && dsb_model_nroots<M>::value == 0 &&
                      !dsb_model_nout<M>::has_out && !dsb_model_has_reset<M>::value) {
            typedef DsbWithSens<M> MS;
            DsbProblemArgs pa;
            int probes = 0;
            std::vector<int32_t> color_full; std::vector<uint8_t> nz_full;
            if (dsb_host::fill_problem_args(*pr, 1, nt, &pa, &probes, &color_full, &nz_full) != DSB_OK) { rc = DSB_BAD_ARG; return; }
            pa.free_running = free_running;
            dsb_host::build_tableau(method, &pa.rk);
            pa.quorum = DSB_DEFAULT_QUORUM; pa.newton_passes = DSB_DEFAULT_NEWTON_PASSES;
            std::vector<double> y0(N), dy0(N), h0(1), ysb((size_t)nt * N), ssb((size_t)nt * NP * N), fin_t(1), fin_h(1);
            std::vector<int32_t> st(DSB_NSTATS), status1(1), fin_order(1), ridx(1), nc(1);
            for (int64_t b = 0; b < B; ++b) {
                DsbBatchBuffers bb{};
                bb.params = params + b * NP; bb.t_eval = t_eval; bb.y0 = y0.data(); bb.dy0 = dy0.data(); bb.h0 = h0.data();
                bb.ys = ysb.data(); bb.ss = ssb.data(); bb.stats = st.data(); bb.status = status1.data();
                bb.fin_t = fin_t.data(); bb.fin_h = fin_h.data(); bb.fin_order = fin_order.data();
                bb.root_idx = ridx.data(); bb.ncols = nc.data();
                for (auto& v : ysb) v = dsb_from_bits(~0ull);
                for (auto& v : ssb) v = dsb_from_bits(~0ull);
                for (auto& v : st) v = 0;
                status1[0] = -1;
                unsigned long long work_counter = 0;
                if (method == DSB_METHOD_BDF) {
                    blockDim.x = BdfLayout<MS>::THREADS;
                    std::vector<double> sens_ws((size_t)BdfLayout<MS>::SDIFF_WORDS * BdfLayout<MS>::THREADS);
                    bb.sens_ws = sens_ws.data();
                    dsb_init_kernel<M>(pa, bb, 1);
                    dsb_bdf_solve_dense_kernel<MS>(pa, bb, &work_counter);
                } else {
                    blockDim.x = SdirkLayout<MS>::THREADS;
                    dsb_init_kernel<M>(pa, bb, pa.rk.order);
                    dsb_sdirk_solve_dense_kernel<MS>(pa, bb, &work_counter);
                }
                for (size_t k = 0; k < ysb.size(); ++k) ys[b * ysb.size() + k] = ysb[k];
                for (size_t k = 0; k < ssb.size(); ++k) sens[b * ssb.size() + k] = ssb[k];
                for (int s = 0; s < DSB_NSTATS; ++s) stats[b * DSB_NSTATS + s] = st[s] + (s == DSB_STAT_RHS_JAC_MULS ? probes : 0);
                status[b] = status1[0];
            }
            rc = DSB_OK;
        } else {
            rc = DSB_ERR;
        }
    }
};

// solve(final_time): the on-chip lane kernels instantiated for DsbRagged<M> (dsb_inst.cu: RaggedLauncher), writing pass with
// the instance's run at offset 0 of a max_cols-column block
struct EmuRaggedCall {
    const dsb_problem* pr; int method;
    const double* params; int64_t B; double final_time; int max_cols;
    double* ts; double* ys; int32_t* ncols; int64_t* stats; int32_t* status; int32_t* root_idx;
    int rc;
    template <class M> void operator()() {
        constexpr int N = M::N, NP = M::NP;
        if constexpr (N <= 16 && !dsb_model_has_reset<M>::value) {
            typedef DsbRagged<M> MR;
            constexpr int NOUT = dsb_model_nout<M>::value;
            DsbProblemArgs pa;
            int probes = 0;
            std::vector<int32_t> color_full; std::vector<uint8_t> nz_full;
            if (dsb_host::fill_problem_args(*pr, 1, 1, &pa, &probes, &color_full, &nz_full) != DSB_OK) { rc = DSB_BAD_ARG; return; }
            pa.free_running = 0; pa.ragged = 2;
            dsb_host::build_tableau(method, &pa.rk);
            pa.quorum = DSB_DEFAULT_QUORUM; pa.newton_passes = DSB_DEFAULT_NEWTON_PASSES;
            std::vector<double> y0(N), dy0(N), h0(1), ysb(NOUT), fin_t(1), fin_h(1);
            std::vector<int32_t> st(DSB_NSTATS), status1(1), fin_order(1), ridx(1), nc(1);
            const int64_t off[2] = {0, 0};
            for (int64_t b = 0; b < B; ++b) {
                DsbBatchBuffers bb{};
                bb.params = params + b * NP; bb.t_eval = &final_time; bb.y0 = y0.data(); bb.dy0 = dy0.data(); bb.h0 = h0.data();
                bb.ys = ysb.data(); bb.ss = nullptr; bb.stats = st.data(); bb.status = status1.data();
                bb.fin_t = fin_t.data(); bb.fin_h = fin_h.data(); bb.fin_order = fin_order.data();
                bb.root_idx = ridx.data(); bb.ncols = nc.data();
                bb.rag_off = off; bb.rag_ts = ts + b * max_cols; bb.rag_ys = ys + b * max_cols * NOUT;
                for (auto& v : st) v = 0;
                status1[0] = -1; ridx[0] = -1; nc[0] = 0;
                unsigned long long work_counter = 0;
                if (method == DSB_METHOD_BDF) {
                    blockDim.x = BdfLayout<MR>::THREADS;
                    dsb_init_kernel<M>(pa, bb, 1);
                    dsb_bdf_solve_dense_kernel<MR>(pa, bb, &work_counter);
                } else {
                    blockDim.x = SdirkLayout<MR>::THREADS;
                    dsb_init_kernel<M>(pa, bb, pa.rk.order);
                    dsb_sdirk_solve_dense_kernel<MR>(pa, bb, &work_counter);
                }
                for (int s = 0; s < DSB_NSTATS; ++s) stats[b * DSB_NSTATS + s] = st[s] + (s == DSB_STAT_RHS_JAC_MULS ? probes : 0);
                status[b] = status1[0]; ncols[b] = nc[0]; root_idx[b] = ridx[0];
            }
            rc = DSB_OK;
        } else {
            rc = DSB_ERR;
        }
    }
};

}  // namespace

// SmemBandLU (dsb_wband_bdf_kernel.cuh) on a dense column-major n x n matrix whose entries outside the band (kl, ku) are
// ignored: band storage, factor, then solve b in place.  mode 0: the fast back substitution (the interchange-free
// forward sweep when the factorisation did not interchange), 1: the exact form.  Returns 0 for a zero pivot, else what
// solve() returned (1, or 2 = a quotient could not be vouched for); *nswaps = row interchanges.
template <int N, int KL, int KU>
static int emu_band_lu_run(const double* A, double* b, int mode, int* nswaps) {
    constexpr int NS = (N + 1) & ~1, KV = KL + KU, LDAB = 2 * KL + KU + 1;
    typedef SmemBandLU<N, NS, KL, KU> BLU;
    std::vector<double> ab((size_t)LDAB * NS, 0.0), rcp(NS, 0.0);
    std::vector<int> piv(NS, 0);
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i)
            if (i - j <= KL && j - i <= KU) ab[(size_t)(KV + i - j) * NS + j] = A[(size_t)j * N + i];
    int swaprow[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int nzero = BLU::factor(ab.data(), piv.data(), rcp.data(), nswaps, swaprow);
    if (nzero) return 0;
    const int nsw = *nswaps;
    if (mode == 1) return BLU::template solve<true, 2>(ab.data(), piv.data(), rcp.data(), b, swaprow, nsw);
    if (mode == 2) return BLU::template solve<false, 2>(ab.data(), piv.data(), rcp.data(), b, swaprow, nsw);      // selects on every row
    return nsw == 0 ? BLU::template solve<false, 0>(ab.data(), piv.data(), rcp.data(), b, swaprow, 0)
         : nsw <= BLU::MAXSW ? BLU::template solve<false, 1>(ab.data(), piv.data(), rcp.data(), b, swaprow, nsw)
                             : BLU::template solve<false, 2>(ab.data(), piv.data(), rcp.data(), b, swaprow, nsw);
}

extern "C" {

void dsb_options_default(dsb_options* o);   // defined below (the same defaults as dsb_capi.cu, problem.rs:132-152)

// kernel: 1 = on-chip lane kernels (n <= 16), 3 = banded lane kernels, 4 = banded warp-per-instance kernel (one lane per warp)
int emu_solve(int model, int method, int kernel, double rtol, const double* atol, int natol, double t0, double h0,
              int use_coloring, const dsb_options* opt, const double* params, int64_t B, const double* t_eval, int nt,
              int free_running, double* ys, int64_t* stats, int32_t* status, double* fin, int32_t* root_idx, int32_t* ncols) {
    dsb_problem pr;
    pr.model = model; pr.n = 0; pr.np = 0; pr.has_mass = 0;
    pr.rtol = rtol; pr.atol.assign(atol, atol + natol); pr.t0 = t0; pr.h0 = h0; pr.use_coloring = use_coloring;
    if (opt) pr.opt = *opt; else dsb_options_default(&pr.opt);
    EmuDims dims{&pr};
    if (!dsb_dispatch_model(model, dims)) return DSB_BAD_ARG;
    if (natol != 1 && natol != pr.n) return DSB_BAD_ARG;
    EmuCall call{&pr, method, kernel, free_running, params, B, t_eval, nt, ys, stats, status, fin, root_idx, ncols, DSB_ERR};
    dsb_dispatch_model(model, call);
    return call.rc;
}

// solve_dense_sensitivities (or, free_running, the step / interpolate / interpolate_sens loop) through the sensitivity
// instantiation of the on-chip BDF lane kernel; ys [B][nt][n], sens [B][nt][np][n].  sens_natol = 0: not in the error test.
int emu_solve_sens(int model, int method, double rtol, const double* atol, int natol, double t0, double h0, const dsb_options* opt,
                   double sens_rtol, const double* sens_atol, int sens_natol, const double* params, int64_t B,
                   const double* t_eval, int nt, int free_running, double* ys, double* sens, int64_t* stats, int32_t* status) {
    dsb_problem pr;
    pr.model = model; pr.n = 0; pr.np = 0; pr.has_mass = 0;
    pr.rtol = rtol; pr.atol.assign(atol, atol + natol); pr.t0 = t0; pr.h0 = h0; pr.use_coloring = 0;
    if (opt) pr.opt = *opt; else dsb_options_default(&pr.opt);
    EmuDims dims{&pr};
    if (!dsb_dispatch_model(model, dims)) return DSB_BAD_ARG;
    if (natol != 1 && natol != pr.n) return DSB_BAD_ARG;
    pr.sens = 1; pr.sens_rtol = sens_rtol; pr.sens_atol.assign(sens_atol, sens_atol + sens_natol);
    EmuSensCall call{&pr, method, free_running, params, B, t_eval, nt, ys, sens, stats, status, DSB_ERR};
    dsb_dispatch_model(model, call);
    return call.rc;
}

// OdeSolverMethod::solve(final_time) through the DsbRagged<M> instantiation of the on-chip lane kernels; instance b's columns
// at ts[b * max_cols ..), ys[b * max_cols * nout ..) -- max_cols must hold the longest run (the emulation does not check)
int emu_solve_ragged(int model, int method, double rtol, const double* atol, int natol, double t0, double h0, const dsb_options* opt,
                     const double* params, int64_t B, double final_time, int max_cols, double* ts, double* ys, int32_t* ncols,
                     int64_t* stats, int32_t* status, int32_t* root_idx) {
    dsb_problem pr;
    pr.model = model; pr.n = 0; pr.np = 0; pr.has_mass = 0;
    pr.rtol = rtol; pr.atol.assign(atol, atol + natol); pr.t0 = t0; pr.h0 = h0; pr.use_coloring = 0;
    if (opt) pr.opt = *opt; else dsb_options_default(&pr.opt);
    EmuDims dims{&pr};
    if (!dsb_dispatch_model(model, dims)) return DSB_BAD_ARG;
    if (natol != 1 && natol != pr.n) return DSB_BAD_ARG;
    EmuRaggedCall call{&pr, method, params, B, final_time, max_cols, ts, ys, ncols, stats, status, root_idx, DSB_ERR};
    dsb_dispatch_model(model, call);
    return call.rc;
}

int emu_smem_band_lu(int n, int kl, int ku, const double* A, double* b, int mode, int* nswaps) {
    if (n == 24 && kl == 1 && ku == 1) return emu_band_lu_run<24, 1, 1>(A, b, mode, nswaps);
    if (n == 24 && kl == 2 && ku == 1) return emu_band_lu_run<24, 2, 1>(A, b, mode, nswaps);
    if (n == 24 && kl == 1 && ku == 2) return emu_band_lu_run<24, 1, 2>(A, b, mode, nswaps);
    if (n == 24 && kl == 2 && ku == 2) return emu_band_lu_run<24, 2, 2>(A, b, mode, nswaps);
    if (n == 7 && kl == 2 && ku == 2) return emu_band_lu_run<7, 2, 2>(A, b, mode, nswaps);
    if (n == 41 && kl == 1 && ku == 1) return emu_band_lu_run<41, 1, 1>(A, b, mode, nswaps);
    return -1;
}

// the product's host-side greedy colouring (dsb_host_setup.h) on a given pattern
int emu_greedy_coloring(const int32_t* rows, const int32_t* cols, int nnz, int n, int32_t* colors) {
    std::vector<std::pair<int, int>> nz;
    for (int k = 0; k < nnz; ++k) nz.push_back({rows[k], cols[k]});
    const std::vector<int> r = dsb_host::greedy_coloring(nz, n);
    for (int j = 0; j < n; ++j) colors[j] = r[j];
    return DSB_OK;
}

void dsb_options_default(dsb_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->max_nonlinear_solver_iterations = 10; o->max_error_test_failures = 40; o->max_nonlinear_solver_failures = 50;
    o->update_jacobian_after_steps = 20; o->update_rhs_jacobian_after_steps = 50;
    o->ic_max_linesearch_iterations = 10; o->ic_max_newton_iterations = 10; o->ic_max_linear_solver_setups = 4;
    o->ic_use_linesearch = 1;
    o->nonlinear_solver_tolerance = 0.2; o->min_timestep = 1e-13;
    o->max_timestep_growth = 2.0; o->min_timestep_growth = 2.0; o->max_timestep_shrink = 0.9; o->min_timestep_shrink = 0.5;
    o->threshold_to_update_jacobian = 0.3; o->threshold_to_update_rhs_jacobian = 0.2;
    o->pi_control_proportional = 0.0; o->pi_control_integral = 0.5;
    o->ic_step_reduction_factor = 0.5; o->ic_armijo_constant = 1e-4;
}

}  // extern "C"
