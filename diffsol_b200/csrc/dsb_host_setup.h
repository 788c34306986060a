// dsb_host_setup.h -- host-side problem setup shared by the C-ABI library (dsb_capi.cu, dsb_inst.cu) and by the
// single-lane host build of the lane kernels that tests/host_emu uses to check kernel logic on machines without a
// GPU: the problem description (what `OdeBuilder::build()` collects, crates/diffsol/src/ode_solver/builder.rs),
// the order-dependent BDF tables, the (E)SDIRK tableaux, sparsity detection + greedy colouring, and the per-column
// metadata of the banded lane kernels.  Plain C++17, no CUDA runtime calls.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

#include "../../include/diffsol_b200.h"
#include "dsb_args.h"
#include "dsb_math.h"
#include "dsb_models.h"

struct dsb_problem {
    int model;
    int n, np, has_mass;
    int nout;                   // rows of a solve_dense column: outputs of the out function, else n
    double rtol;
    std::vector<double> atol;
    double t0, h0;
    int use_coloring;
    dsb_options opt;
    int sens = 0;               // forward sensitivities (dsb_problem_set_sensitivities)
    double sens_rtol = 0.0;
    std::vector<double> sens_atol;   // empty: not in the error test; 1 entry broadcasts
};

namespace dsb_host {

// bdf.rs:253-276 (kappa, gamma, alpha, error_const2), bdf.rs:433-463 (U = R(order, 1)),
// convergence.rs:36-42 (eta resets), line_search.rs:126 (steptol).  Host code of this file is compiled
// with -ffp-contract=off so these are the same doubles the oracle computes.
inline void build_tables(DsbBdfTables* tb, const dsb_options* opt = nullptr) {
    const double kappa[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
    tb->alpha[0] = 0.0; tb->gamma[0] = 0.0; tb->error_const2[0] = 1.0;
    for (int i = 1; i <= DSB_MAX_ORDER; ++i) {
        const double i_t = (double)i;
        const double one_over_i = 1.0 / i_t;
        const double one_over_i_plus_one = 1.0 / (i_t + 1.0);
        tb->gamma[i] = tb->gamma[i - 1] + one_over_i;
        tb->alpha[i] = 1.0 / ((1.0 - kappa[i]) * tb->gamma[i]);
        const double e = kappa[i] * tb->gamma[i] + one_over_i_plus_one;
        tb->error_const2[i] = e * e;
    }
    std::memset(tb->u, 0, sizeof(tb->u));
    for (int order = 1; order <= DSB_MAX_ORDER; ++order) {
        const int nr = order + 1;
        double* r = tb->u[order];
        for (int j = 0; j < nr; ++j) r[j * nr] = 1.0;
        for (int j = 1; j < nr; ++j) {
            const double j_t = (double)j;
            for (int i = 1; i < nr; ++i) {
                const double i_t = (double)i;
                const int idx = j * nr + i;
                r[idx] = r[idx - 1] * (i_t - 1.0 - 1.0 * j_t) / i_t;
            }
        }
    }
    tb->eta_reset = dsb_pow(20.0, 1.25);
    tb->eta_reset_timestep = dsb_pow(100.0, 1.25);
    tb->ic_steptol = dsb_pow(std::numeric_limits<double>::epsilon(), 2.0 / 3.0);
    for (int k = 0; k < 32; ++k) {
        tb->inv_int[k] = k > 0 ? 1.0 / (double)k : 0.0;
        tb->safety[k] = 0.0;
    }
    for (int o = 0; o < DSB_MAX_ORDER + 2; ++o) { tb->pi_ki[o] = 0.0; tb->pi_kp[o] = 0.0; }
    if (opt) {
        const double maxiter = (double)opt->max_nonlinear_solver_iterations;
        for (int k = 0; k < 32; ++k) tb->safety[k] = 0.9 * (2.0 * maxiter + 1.0) / (2.0 * maxiter + (double)k);
        for (int o = 1; o < DSB_MAX_ORDER + 2; ++o) {
            tb->pi_ki[o] = opt->pi_control_integral / (double)o;
            tb->pi_kp[o] = opt->pi_control_proportional / (double)o;
        }
    }
}

// ode_solver/tableau.rs:41-97 (tr_bdf2) and :101-159 (esdirk34); same expressions as the reference
inline void build_tableau(int method, DsbSdirkTableau* t) {
    std::memset(t, 0, sizeof(*t));
    if (method == DSB_METHOD_TR_BDF2) {
        t->s = 3; t->order = 2; t->has_beta = 1;
        const double gamma = 2.0 - std::sqrt(2.0);
        const double d = gamma / 2.0;
        const double w = std::sqrt(2.0) / 4.0;
        const double a[9] = {0.0, d, w, 0.0, d, w, 0.0, 0.0, d};
        for (int i = 0; i < 9; ++i) t->a[i] = a[i];
        t->b[0] = w; t->b[1] = w; t->b[2] = d;
        const double b_hat[3] = {(1.0 - w) / 3.0, (3.0 * w + 1.0) / 3.0, d / 3.0};
        for (int i = 0; i < 3; ++i) t->d[i] = t->b[i] - b_hat[i];
        const double beta[6] = {2.0 * w, 2.0 * w, gamma - 1.0, -w, -w, 2.0 * w};
        for (int i = 0; i < 6; ++i) t->beta[i] = beta[i];
        t->c[0] = 0.0; t->c[1] = gamma; t->c[2] = 1.0;
    } else if (method == DSB_METHOD_ESDIRK34) {
        t->s = 4; t->order = 3; t->has_beta = 0;
        const double g = 0.435866521508459;
        const double a[16] = {0.0, g, 0.1407377747247062, 0.102399400619911,
                              0.0, g, -0.1083655513813208, -0.3768784522555561,
                              0.0, 0.0, g, 0.8386125301271861,
                              0.0, 0.0, 0.0, g};
        for (int i = 0; i < 16; ++i) t->a[i] = a[i];
        for (int j = 0; j < 4; ++j) t->b[j] = a[j * 4 + 3];
        const double c[4] = {0.0, 0.871733043016918, 0.4682387448518444, 1.0};
        const double d[4] = {-0.05462549724041394, -0.49420889362599496, 0.22193449973506466, 0.32689989113134427};
        for (int i = 0; i < 4; ++i) { t->c[i] = c[i]; t->d[i] = d[i]; }
    }
}

// coloring.rs:27-47 (nonzeros2graph: columns that share a row are adjacent) + greedy_coloring.rs:14-34
// (color_graph_greedy): colour (1-based) of every column
inline std::vector<int> greedy_coloring(const std::vector<std::pair<int, int>>& non_zeros, int N) {
    std::vector<std::vector<int>> cols_by_rows(N), adj(N);
    for (auto& ij : non_zeros) cols_by_rows[ij.first].push_back(ij.second);
    for (auto& ij : non_zeros)
        for (int next_col : cols_by_rows[ij.first])
            if (next_col < ij.second) { adj[ij.second].push_back(next_col); adj[next_col].push_back(ij.second); }
    std::vector<int> result(N, 0);
    if (N > 0) result[0] = 1;
    std::vector<char> available(N, 0);
    for (int ii = 1; ii < N; ++ii) {
        for (int j : adj[ii]) if (result[j] != 0) available[result[j] - 1] = 1;
        for (int i = 0; i < N; ++i) if (!available[i]) { result[ii] = i + 1; break; }
        std::fill(available.begin(), available.end(), 0);
    }
    return result;
}

// jacobian/mod.rs:16-48 (NaN probe), coloring.rs:27-47 (graph), greedy_coloring.rs:14-34.
// The pattern is a property of the equations, not of the instance ("assume every batch has the same
// non-zeros", jacobian/mod.rs:32), so it is found once on the host with the model's own functor.
struct ColoringOf {
    const dsb_problem* pr; DsbProblemArgs* pa; int* probes;
    std::vector<int32_t>* color_full; std::vector<uint8_t>* nz_full;      // any n: for the block-per-instance path
    template <class M> void operator()() {
        constexpr int N = M::N;
        constexpr int NP = M::NP;
        double p[NP > 0 ? NP : 1];
        for (int j = 0; j < NP; ++j) p[j] = 1.0;
        std::vector<double> y0v(N), vv(N), colv(N);
        double* y0 = y0v.data(); double* v = vv.data(); double* col = colv.data();
        M::init(p, pr->t0, y0);
        std::vector<std::pair<int, int>> non_zeros;
        for (int i = 0; i < N; ++i) { v[i] = 0.0; col[i] = 0.0; }
        for (int j = 0; j < N; ++j) {
            v[j] = std::numeric_limits<double>::quiet_NaN();
            M::jac_mul(y0, p, pr->t0, v, col);
            for (int i = 0; i < N; ++i) if (std::isnan(col[i])) non_zeros.push_back({i, j});
            for (int i = 0; i < N; ++i) col[i] = 0.0;
            v[j] = 0.0;
        }
        *probes = N;
        const std::vector<int> result = greedy_coloring(non_zeros, N);
        int max_color = 0;
        for (int c : result) if (c > max_color) max_color = c;
        pa->ncolors = max_color;
        if (N <= DSB_MAX_STATES) {
            for (int j = 0; j < N; ++j) { pa->color_of_col[j] = result[j] - 1; pa->nz_rows_of_col[j] = 0; }
            for (auto& ij : non_zeros) pa->nz_rows_of_col[ij.second] |= (1ull << ij.first);
        }
        // dense form: colour of every column (-1: the column has no non-zero and is never seeded), pattern bytes
        color_full->assign(N, -1);
        nz_full->assign((size_t)N * N, 0);
        for (auto& ij : non_zeros) { (*nz_full)[(size_t)ij.second * N + ij.first] = 1; (*color_full)[ij.second] = result[ij.second] - 1; }
    }
};

// set by dsb_capi.cu when a model plugin is loaded: pattern + colouring of a model that is not in the built-in registry
typedef int (*plugin_coloring_fn)(const dsb_problem&, DsbProblemArgs*, int*, std::vector<int32_t>*, std::vector<uint8_t>*);
inline plugin_coloring_fn& plugin_coloring() { static plugin_coloring_fn f = nullptr; return f; }

inline int fill_problem_args(const dsb_problem& pr, int64_t B, int nt, DsbProblemArgs* pa, int* probes,
                      std::vector<int32_t>* color_full, std::vector<uint8_t>* nz_full) {
    std::memset(pa, 0, sizeof(*pa));
    pa->nbatch = B; pa->nt = nt;
    pa->rtol = pr.rtol; pa->t0 = pr.t0; pa->h0 = pr.h0;
    for (int i = 0; i < pr.n && i < DSB_MAX_STATES; ++i) pa->atol[i] = pr.atol.size() == 1 ? pr.atol[0] : pr.atol[i];
    pa->opt = pr.opt;
    build_tables(&pa->tab, &pr.opt);
    pa->use_coloring = pr.use_coloring;
    pa->sens = pr.sens;
    pa->sens_error_control = pr.sens && !pr.sens_atol.empty();
    pa->sens_rtol = pr.sens_rtol;
    for (int i = 0; i < pr.n && i < DSB_MAX_STATES && !pr.sens_atol.empty(); ++i)
        pa->sens_atol[i] = pr.sens_atol.size() == 1 ? pr.sens_atol[0] : pr.sens_atol[i];
    *probes = 0;
    if (pr.use_coloring) {
        ColoringOf f{&pr, pa, probes, color_full, nz_full};
        if (pr.model >= DSB_MODEL_PLUGIN_ID0) {
            if (!plugin_coloring() || plugin_coloring()(pr, pa, probes, color_full, nz_full) != DSB_OK) return DSB_BAD_ARG;
        } else if (!dsb_dispatch_model(pr.model, f)) return DSB_BAD_ARG;
    } else {
        // dense assembly (op/nonlinear_op.rs:211-220) expressed as one colour per column with a full pattern, so
        // that the lane kernels carry a single assembly loop (dsb_lane.cuh:lane_jacobian_to; the banded kernel
        // gets the same tables for any n as a device array, dsb_inst.cu:BandLauncher)
        pa->ncolors = pr.n;
        for (int j = 0; j < pr.n && j < DSB_MAX_STATES; ++j) {
            pa->color_of_col[j] = j;
            pa->nz_rows_of_col[j] = pr.n >= 64 ? ~0ull : ((1ull << pr.n) - 1ull);
        }
    }
    return DSB_OK;
}

// Per-column metadata of df/dy for the banded lane kernels (dsb_band_bdf_kernel.cuh: DsbBandMeta): sparsity pattern by
// NaN probe (jacobian/mod.rs:16-48) -- the declared band (kl, ku) must cover it, else false -- packed per column as
// colour | (in-band pattern << 16).  With colouring the colours come from the greedy colouring (`color_host`, -1 =
// the column has no non-zero); without, the dense assembly is expressed as one colour per column that stores every
// in-band entry.
template <class M>
inline bool band_column_meta(double t0, bool use_coloring, const int32_t* color_host, int kl, int ku, std::vector<int32_t>* colmeta) {
    constexpr int N = M::N;
    colmeta->assign(N, 0);
    double p[M::NP > 0 ? M::NP : 1];
    for (int j = 0; j < M::NP; ++j) p[j] = 1.0;
    std::vector<double> y0(N), v(N, 0.0), col(N, 0.0);
    M::init(p, t0, y0.data());
    for (int j = 0; j < N; ++j) {
        v[j] = std::numeric_limits<double>::quiet_NaN();
        M::jac_mul(y0.data(), p, t0, v.data(), col.data());
        int32_t mask = 0;
        for (int i = 0; i < N; ++i) {
            if (!std::isnan(col[i])) continue;
            if (i - j > kl || j - i > ku) return false;
            mask |= 1 << (ku + i - j);
        }
        for (int i = 0; i < N; ++i) col[i] = 0.0;
        v[j] = 0.0;
        if (use_coloring) {
            const int32_t colour = color_host[j] < 0 ? 0xffff : color_host[j];
            (*colmeta)[j] = colour | (mask << 16);
        } else {
            (*colmeta)[j] = j | (((1 << (kl + ku + 1)) - 1) << 16);
        }
    }
    return true;
}

}  // namespace dsb_host
