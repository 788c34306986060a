// dsb_roots.cuh -- per-lane restatement of the reference's event detection, shared by every lane kernel.
//
//   RootFinder::{init, check_root}   crates/diffsol/src/nonlinear_solver/root.rs:12-160 (the modified secant / Illinois
//                                    iteration of SUNDIALS' rootfinding)
//   Vector::root_finding             crates/diffsol-la/src/vector/nalgebra_serial.rs:484-504
// called after every accepted step by Bdf::step (ode_solver/bdf.rs:1566-1579) and Rk::step_accepted
// (ode_solver/runge_kutta.rs:935-948).  Storage-agnostic: the caller passes two evaluators,
//   end(g)        g = root_fn(state.y, state.t)
//   at(t_mid, g)  g = root_fn(interpolate(t_mid), t_mid)
// so the on-chip kernels (state in registers / shared memory) and the banded kernels (state in global memory) use
// the same iteration.  Only instantiated for equation sets that declare NROOTS > 0.
#pragma once
#include "dsb_lane.cuh"

template <int NR, class DIV>
struct LaneRootFinder {
    double g0[NR];          // root function at the lower end of the search interval
    double t0;

    static DSB_DEV void root_finding(const double (&ga)[NR], const double (&gb)[NR], bool& found, int& imax) {
        double max_frac = 0.0;
        imax = -1; found = false;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if (gb[r] == 0.0) found = true;
            if (ga[r] * gb[r] < 0.0) {
                const double frac = dsb_abs(DIV::div(gb[r], gb[r] - ga[r]));
                if (frac > max_frac) { max_frac = frac; imax = r; }
            }
        }
    }
    static DSB_DEV double pick(const double (&g)[NR], int k) {
        double v = g[0];
#pragma unroll
        for (int r = 1; r < NR; ++r) if (k == r) v = g[r];
        return v;
    }

    // true <=> a root lies in (t0, t]: t_root and the index of the root function are set.  Otherwise the interval's
    // lower end moves to t.
    template <class FE, class FA>
    DSB_DEV bool check_root(const double t, FE&& end, FA&& at, double& t_root, int& root_idx) {
        const double eps = 2.220446049250313e-16;
        double g_end[NR], g1[NR], gmid[NR];
        end(g_end);
#pragma unroll
        for (int r = 0; r < NR; ++r) g1[r] = g_end[r];
        bool rootfnd; int imax;
        root_finding(g0, g1, rootfnd, imax);
        t_root = t;
        if (imax < 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) g0[r] = g1[r];
            t0 = t;
            if (rootfnd) {                              // find_zero_index (root.rs:44-58): smallest |g|, first one on ties
                int min_idx = 0; double min_val = dsb_abs(g0[0]);
#pragma unroll
                for (int r = 1; r < NR; ++r) { const double v = dsb_abs(g0[r]); if (v < min_val) { min_val = v; min_idx = r; } }
                root_idx = min_idx;
                return true;
            }
            return false;
        }
        double alpha = 1.0;
        bool sc0 = false, sc1 = true;
        int it = 0;
        double t1 = t, tl = t0;
        const double tol = 100.0 * eps * (dsb_abs(t1) + dsb_abs(t1 - tl));
        bool done = false;
        while (!done && dsb_abs(t1 - tl) > tol) {
            const double g1_val = pick(g1, imax), g0_val = pick(g0, imax);
            double t_mid = t1 - DIV::div((t1 - tl) * g1_val, g1_val - alpha * g0_val);
            if (dsb_abs(t_mid - tl) < 0.5 * tol) {
                const double fracint = DIV::div(dsb_abs(t1 - tl), tol);
                const double fracsub = fracint > 5.0 ? 0.1 : DIV::div(0.5, fracint);
                t_mid = tl + fracsub * (t1 - tl);
            }
            if (dsb_abs(t1 - t_mid) < 0.5 * tol) {
                const double fracint = DIV::div(dsb_abs(t1 - tl), tol);
                const double fracsub = fracint > 5.0 ? 0.1 : DIV::div(0.5, fracint);
                t_mid = t1 - fracsub * (t1 - tl);
            }
            at(t_mid, gmid);
            bool rf; int im;
            root_finding(g0, gmid, rf, im);
            const bool lower = im >= 0;
            if (lower) {
                t1 = t_mid; imax = im;
#pragma unroll
                for (int r = 0; r < NR; ++r) g1[r] = gmid[r];
            } else if (rf) {
                t_root = t_mid; done = true;
            } else {
                tl = t_mid;
#pragma unroll
                for (int r = 0; r < NR; ++r) g0[r] = gmid[r];
            }
            if (!done) {
                if ((it & 1) == 0) sc0 = lower; else sc1 = lower;
                if (it >= 2) alpha = (sc0 != sc1) ? 1.0 : (sc0 ? 0.5 * alpha : 2.0 * alpha);
                ++it;
            }
        }
        if (!done) t_root = t1;
#pragma unroll
        for (int r = 0; r < NR; ++r) g0[r] = g_end[r];          // root_fn(y, t) again, into g0
        root_idx = imax;
        return true;
    }
};
