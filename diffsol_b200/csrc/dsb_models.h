// dsb_models.h -- the user-equation side of the boundary: rhs / jac_mul / mass / init functors.
//
// These play the role of the Rust closures a diffsol user hands to
// `OdeBuilder::rhs_implicit(f, jac_mul).mass(m).init(i, n)` (crates/diffsol/src/ode_solver/builder.rs:192-200):
//   rhs     (x, p, t, y)        y = f(x, p, t)                 Fn(&V,&V,T,&mut V)
//   jac_mul (x, p, t, v, y)     y = df/dx(x, p, t) . v         Fn(&V,&V,T,&V,&mut V)
//   mass    (x, p, t, beta, y)  y = M x + beta y               Fn(&V,&V,T,T,&mut V)
//   init    (p, t, y)           y = y0(p, t)                   Fn(&V,T,&mut V)
// They are written once as __host__ __device__ code so the sm_100a kernels and the CPU oracle
// integrate *the same equations with the same floating-point expression trees*; the expression
// order follows the reference closures literally (file:line cited on each model).
#pragma once
#include <type_traits>

#include "dsb_math.h"

enum dsb_model_id {
    DSB_MODEL_EXP_DECAY = 0,            // n=2  np=2   test_models/exponential_decay.rs:14-21,54-61,75-84
    DSB_MODEL_EXP_DECAY_ALGEBRAIC = 1,  // n=3  np=1   test_models/exponential_decay_with_algebraic.rs:18-23,59-71,94-106,122-126
    DSB_MODEL_ROBERTSON_DAE = 2,        // n=3  np=3   test_models/robertson.rs:59-90
    DSB_MODEL_ROBERTSON_ODE = 3,        // n=3  np=3   test_models/robertson_ode.rs:46-104 (ngroups=1)
    DSB_MODEL_ROBERTSON_ODE_G3 = 4,     // n=9  np=3   same, ngroups=3
    DSB_MODEL_DYDT_Y2 = 5,              // n=10 np=0   test_models/dydt_y2.rs:9-19
    DSB_MODEL_GAUSSIAN_DECAY = 6,       // n=10 np=10  test_models/gaussian_decay.rs:12-23
    DSB_MODEL_VAN_DER_POL = 7,          // n=2  np=1   (not in the reference; BASELINE.json config 3)
    DSB_MODEL_VAN_DER_POL_SCALED = 8,   // n=2  np=2   the same in scaled time tau = t / T, p = [mu, T]
    DSB_MODEL_HEAT1D_DAE_256 = 9,       // n=256 np=3  1-D heat equation, boundary rows algebraic (BASELINE.json config 4)
    DSB_MODEL_HEAT1D_DAE_32 = 10,       // n=32  np=3  the same on a coarse grid (test size)
    DSB_MODEL_SPM = 11,                 // n=42  np=1  single-particle battery model, book/src/primer/src/spm.ds (BASELINE config 5)
    DSB_MODEL_SPM99 = 12,               // n=200 np=1  the same model on 99 radial cells per particle (not in the reference)
    DSB_MODEL_EXP_DECAY_ROOT = 13,      // n=2  np=2   exp_decay with the root y[0] - 0.6 (test_models/exponential_decay.rs:370-390)
    DSB_MODEL_SPM_STOP = 14,            // n=42 np=1   spm with the model text's out (terminal voltage) and stop (voltage leaves [3.105, 4.1] V) functions
    DSB_MODEL_SPM99_STOP = 15,          // n=200 np=1  the same on 99 radial cells per particle
    DSB_MODEL_HEAT1D_DAE_32_BC = 16,    // n=32 np=3   heat1d_dae_32 with warm boundaries 0 = u - height/4: INCONSISTENT initial values
    DSB_MODEL_EXP_DECAY_RESET = 17,     // n=2  np=2   exp_decay with the roots y[0] - 0.6, y[0] - 0.3 and the reset y -> 0.4 (exponential_decay.rs:818-880)
    DSB_MODEL_HEAT2D_10 = 18,           // n=100 np=0  2-D heat equation on a 10 x 10 grid, boundary rows algebraic (test_models/heat2d.rs), states only
    DSB_MODEL_BALL_BOUNCE = 19,         // n=2  np=3   bouncing ball x' = v, v' = -g with the root x and the reset v -> -e v (ode_solver/mod.rs:1001-1080)
    DSB_MODEL_EXP_DECAY_TWO_ROOTS = 20, // n=2  np=2   exp_decay with the roots y[0] - 0.6, y[0] - 0.3 and no reset (exponential_decay.rs:827-832, 890-912)
    DSB_MODEL_SPM_CYCLE = 21,           // n=42 np=1   spm_stop with a reset: at a voltage cut-off the cell goes back to its initial (charged) state
    DSB_MODEL_EXP_DECAY_ALGEBRAIC_RESET = 22,  // n=3 np=2  exp_decay_algebraic with p = [k, y0], the roots y[0] - 0.6, y[0] - 2 and the reset y -> y + 2 (a DAE with a reset: apply_reset_with_mass)
    DSB_MODEL_COUNT
};

// dy/dt = -k y, p = [k, y0]
struct ModelExpDecay {
    static constexpr int N = 2, NP = 2;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        const double mk = -p[0];
        for (int i = 0; i < N; ++i) y[i] = x[i] * mk;
    }
    DSB_HD static void jac_mul(const double*, const double* p, double, const double* v, double* y) {
        const double mk = -p[0];
        for (int i = 0; i < N; ++i) y[i] = v[i] * mk;
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double* p, double, double* y) {
        for (int i = 0; i < N; ++i) y[i] = p[1];
    }
    // forward sensitivities: exponential_decay_sens / exponential_decay_init_sens (test_models/exponential_decay.rs:
    // f_p v = x * (-v[0]), (d y0 / d p) v = [v[1], v[1]]), the problem of exponential_decay_problem_sens (:703-742)
    static constexpr bool HAS_SENS = true;
    DSB_HD static void sens_mul(const double* x, const double*, double, const double* v, double* y) {
        const double mv = -v[0];
        for (int i = 0; i < N; ++i) y[i] = x[i] * mv;
    }
    DSB_HD static void init_sens(const double*, double, const double* v, double* y) {
        for (int i = 0; i < N; ++i) y[i] = v[1];
    }
};

// The same equations with the root function of the reference's event tests
// (test_models/exponential_decay.rs:102-104, 370-390): g = y[0] - 0.6; the integration stops at the first root.
struct ModelExpDecayRoot : ModelExpDecay {
    static constexpr int NROOTS = 1;
    DSB_HD static void root(const double* x, const double*, double, double* g) { g[0] = x[0] - 0.6; }
};

// dy/dt = -a y ; 0 = z - y ; p = [a]; inconsistent IC [1,1,0]
// The reference's reset test problem (test_models/exponential_decay.rs:818-880, exponential_decay_with_reset_problem):
// roots g0 = y[0] - 0.6 and g1 = y[0] - 0.3, reset y -> [0.4, 0.4]: with a reset function a root does not end
// solve_dense -- the state is reset and the integration goes on to the last t_eval (method.rs:783-797).
struct ModelExpDecayReset : ModelExpDecay {
    static constexpr int NROOTS = 2;
    static constexpr bool HAS_RESET = true;
    DSB_HD static void root(const double* x, const double*, double, double* g) { g[0] = x[0] - 0.6; g[1] = x[0] - 0.3; }
    DSB_HD static void reset(const double*, const double*, double, double* y) { for (int i = 0; i < N; ++i) y[i] = 0.4; }
};

// The reference's root-index test problem (test_models/exponential_decay.rs:890-912, exponential_decay_with_two_roots_problem):
// the same two root functions without a reset; the solve ends at the first root and reports which one fired.
struct ModelExpDecayTwoRoots : ModelExpDecay {
    static constexpr int NROOTS = 2;
    DSB_HD static void root(const double* x, const double*, double, double* g) { g[0] = x[0] - 0.6; g[1] = x[0] - 0.3; }
};

// The reference's bouncing-ball test (ode_solver/mod.rs:1001-1080: DiffSL text `g { 9.81 } h { 10.0 } u_i { x = h, v = 0 }
// F_i { v, -g } stop { x }`, restitution e = 0.8 applied by the test at the root: v <- -e v, x <- max(x, eps)), with
// g, h and e as the instance's parameters p = [g, h, e] and the test's state update as the reset function.
struct ModelBallBounce {
    static constexpr int N = 2, NP = 3;
    static constexpr bool HAS_MASS = false;
    static constexpr int NROOTS = 1;
    static constexpr bool HAS_RESET = true;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) { y[0] = x[1]; y[1] = -p[0]; }
    DSB_HD static void jac_mul(const double*, const double*, double, const double* v, double* y) { y[0] = v[1]; y[1] = 0.0; }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double* p, double, double* y) { y[0] = p[1]; y[1] = 0.0; }
    DSB_HD static void root(const double* x, const double*, double, double* g) { g[0] = x[0]; }
    DSB_HD static void reset(const double* x, const double* p, double, double* y) {
        const double eps = 2.220446049250313e-16;
        y[0] = x[0] > eps ? x[0] : eps;
        y[1] = x[1] * -p[2];
    }
};

struct ModelExpDecayAlgebraic {
    static constexpr int N = 3, NP = 1;
    static constexpr bool HAS_MASS = true;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        const double ma = -p[0];
        for (int i = 0; i < N; ++i) y[i] = x[i] * ma;
        y[N - 1] = x[N - 1] - x[N - 2];
    }
    DSB_HD static void jac_mul(const double*, const double* p, double, const double* v, double* y) {
        const double ma = -p[0];
        for (int i = 0; i < N; ++i) y[i] = v[i] * ma;
        y[N - 1] = v[N - 1] - v[N - 2];
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        const double yn = beta * y[N - 1];
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
        y[N - 1] = yn;
    }
    DSB_HD static void init(const double*, double, double* y) {
        y[0] = 1.0; y[1] = 1.0; y[2] = 0.0;
    }
    // forward sensitivities: exponential_decay_with_algebraic_sens / _init_sens
    // (test_models/exponential_decay_with_algebraic.rs:32-43, 128-135), the problem of ..._problem_sens (:418-453)
    static constexpr bool HAS_SENS = true;
    DSB_HD static void sens_mul(const double* x, const double*, double, const double* v, double* y) {
        const double mv = -v[0];
        for (int i = 0; i < N; ++i) y[i] = x[i] * mv;
        y[N - 1] = 0.0;
    }
    DSB_HD static void init_sens(const double*, double, const double*, double* y) { for (int i = 0; i < N; ++i) y[i] = 0.0; }
};

// The DAE of the reference's reset-with-mass problem (test_models/exponential_decay_with_algebraic.rs:501-560,
// exponential_decay_with_algebraic_with_reset_problem_sens, without its sensitivities): p = [k, y0], every state starts at y0,
// roots y[0] - 0.6 and y[0] - 2.0 (exponential_decay_root_0_6_and_2_0), reset y -> y + 2 (exponential_decay_reset_y_plus_2).
// A reset on a DAE goes through state.apply_reset_with_mass (ode_solver/state.rs:279-306): reset, then set_consistent.
struct ModelExpDecayAlgebraicReset : ModelExpDecayAlgebraic {
    static constexpr int NP = 2;
    static constexpr int NROOTS = 2;
    static constexpr bool HAS_RESET = true;
    static constexpr bool HAS_SENS = false;
    DSB_HD static void init(const double* p, double, double* y) { for (int i = 0; i < N; ++i) y[i] = p[1]; }
    DSB_HD static void root(const double* x, const double*, double, double* g) { g[0] = x[0] - 0.6; g[1] = x[0] - 2.0; }
    DSB_HD static void reset(const double* x, const double*, double, double* y) { for (int i = 0; i < N; ++i) y[i] = x[i] + 2.0; }
};

// Robertson chemical kinetics as an index-1 DAE, p = [k1, k2, k3]
struct ModelRobertsonDae {
    static constexpr int N = 3, NP = 3;
    static constexpr bool HAS_MASS = true;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        y[0] = -p[0] * x[0] + p[1] * x[1] * x[2];
        y[1] = p[0] * x[0] - p[1] * x[1] * x[2] - p[2] * x[1] * x[1];
        y[2] = x[0] + x[1] + x[2] - 1.0;
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) {
        y[0] = -p[0] * v[0] + p[1] * v[1] * x[2] + p[1] * x[1] * v[2];
        y[1] = p[0] * v[0] - p[1] * v[1] * x[2] - p[1] * x[1] * v[2] - 2.0 * p[2] * x[1] * v[1];
        y[2] = v[0] + v[1] + v[2];
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        y[0] = x[0] + beta * y[0];
        y[1] = x[1] + beta * y[1];
        y[2] = beta * y[2];
    }
    DSB_HD static void init(const double*, double, double* y) {
        y[0] = 1.0; y[1] = 0.0; y[2] = 0.0;
    }
    // forward sensitivities: robertson_sens_mul / robertson_init_sens (test_models/robertson.rs:73-77, 91-93), the problem of
    // robertson_sens (:151-201): a DAE, so the sensitivities are made consistent first (state.rs:167-238); the run of
    // bdf.rs:2248-2271 (28 failed Newton solves) pins the oracle and is reproduced by the kernel.
    static constexpr bool HAS_SENS = true;
    DSB_HD static void sens_mul(const double* x, const double*, double, const double* v, double* y) {
        y[0] = -v[0] * x[0] + v[1] * x[1] * x[2];
        y[1] = v[0] * x[0] - v[1] * x[1] * x[2] - v[2] * x[1] * x[1];
        y[2] = 0.0;
    }
    DSB_HD static void init_sens(const double*, double, const double*, double* y) { y[0] = 0.0; y[1] = 0.0; y[2] = 0.0; }
};

// Robertson as a pure ODE, NG decoupled copies (robertson_ode.rs `ngroups`)
template <int NG>
struct ModelRobertsonOde {
    static constexpr int N = 3 * NG, NP = 3;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        for (int ig = 0; ig < NG; ++ig) {
            const int i = ig * 3;
            y[i] = -p[0] * x[i] + p[1] * x[i + 1] * x[i + 2];
            y[i + 1] = p[0] * x[i] - p[1] * x[i + 1] * x[i + 2] - p[2] * x[i + 1] * x[i + 1];
            y[i + 2] = p[2] * x[i + 1] * x[i + 1];
        }
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) {
        for (int ig = 0; ig < NG; ++ig) {
            const int i = ig * 3;
            y[i] = -p[0] * v[i] + p[1] * v[i + 1] * x[i + 2] + p[1] * x[i + 1] * v[i + 2];
            y[i + 1] = p[0] * v[i] - p[1] * v[i + 1] * x[i + 2] - p[1] * x[i + 1] * v[i + 2]
                       - 2.0 * p[2] * x[i + 1] * v[i + 1];
            y[i + 2] = 2.0 * p[2] * x[i + 1] * v[i + 1];
        }
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double*, double, double* y) {
        for (int ig = 0; ig < NG; ++ig) { y[3 * ig] = 1.0; y[3 * ig + 1] = 0.0; y[3 * ig + 2] = 0.0; }
    }
    // forward sensitivities: the closures of robertson_ode_with_sens (test_models/robertson_ode_with_sens.rs:38-50)
    static constexpr bool HAS_SENS = true;
    DSB_HD static void sens_mul(const double* x, const double*, double, const double* v, double* y) {
        for (int ig = 0; ig < NG; ++ig) {
            const int i = ig * 3;
            y[i] = -v[0] * x[i] + v[1] * x[i + 1] * x[i + 2];
            y[i + 1] = v[0] * x[i] - v[1] * x[i + 1] * x[i + 2] - v[2] * x[i + 1] * x[i + 1];
            y[i + 2] = v[2] * x[i + 1] * x[i + 1];
        }
    }
    DSB_HD static void init_sens(const double*, double, const double*, double* y) {
        for (int i = 0; i < N; ++i) y[i] = 0.0;
    }
};

// dy/dt = y^2, y0 = -200
template <int NS>
struct ModelDydtY2 {
    static constexpr int N = NS, NP = 0;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double*, double, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] * x[i];
    }
    DSB_HD static void jac_mul(const double* x, const double*, double, const double* v, double* y) {
        for (int i = 0; i < N; ++i) y[i] = v[i] * x[i] * 2.0;
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double*, double, double* y) {
        for (int i = 0; i < N; ++i) y[i] = -200.0;
    }
};

// dy/dt = -a t y, p = [a_0..a_{n-1}]
template <int NS>
struct ModelGaussianDecay {
    static constexpr int N = NS, NP = NS;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) {
        const double mt = -t;
        for (int i = 0; i < N; ++i) y[i] = x[i] * p[i] * mt;
    }
    DSB_HD static void jac_mul(const double*, const double* p, double t, const double* v, double* y) {
        const double mt = -t;
        for (int i = 0; i < N; ++i) y[i] = v[i] * p[i] * mt;
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double*, double, double* y) {
        for (int i = 0; i < N; ++i) y[i] = 1.0;
    }
};

// Van der Pol oscillator y1' = y2, y2' = mu (1 - y1^2) y2 - y1, p = [mu], y0 = [2, 0]
struct ModelVanDerPol {
    static constexpr int N = 2, NP = 1;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        y[0] = x[1];
        y[1] = p[0] * (1.0 - x[0] * x[0]) * x[1] - x[0];
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) {
        y[0] = v[1];
        y[1] = p[0] * (-2.0 * x[0] * v[0]) * x[1] + p[0] * (1.0 - x[0] * x[0]) * v[1] - v[0];
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double*, double, double* y) {
        y[0] = 2.0; y[1] = 0.0;
    }
};

// Van der Pol in scaled time tau = t / T: dy/dtau = T f(y; mu), p = [mu, T].  Lets a batch whose instances
// need different end times T_i (BASELINE config 3: T = max(20, 2 mu)) share one t_eval grid on [0, 1].
struct ModelVanDerPolScaled {
    static constexpr int N = 2, NP = 2;
    static constexpr bool HAS_MASS = false;
    DSB_HD static void rhs(const double* x, const double* p, double, double* y) {
        y[0] = p[1] * x[1];
        y[1] = p[1] * (p[0] * (1.0 - x[0] * x[0]) * x[1] - x[0]);
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double, const double* v, double* y) {
        y[0] = p[1] * v[1];
        y[1] = p[1] * (p[0] * (-2.0 * x[0] * v[0]) * x[1] + p[0] * (1.0 - x[0] * x[0]) * v[1] - v[0]);
    }
    DSB_HD static void mass(const double* x, const double*, double, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
    }
    DSB_HD static void init(const double*, double, double* y) {
        y[0] = 2.0; y[1] = 0.0;
    }
};

// 1-D heat equation u_t = D u_xx on [0, 1] with NS grid points, D = 0.1 (examples/pde-heat/src/main.rs:18),
// as an index-1 DAE in the style of test_models/heat2d.rs: the two boundary rows are algebraic (0 = u - 0,
// M = diag(0, 1, .., 1, 0)).  p = [height, x_left, x_right]: the initial condition is a plateau of that height
// on [x_left, x_right] (BASELINE.json config 4: an initial-condition sweep).  Written component-wise
// (`*_i`) so that the block-cooperative kernels evaluate one component per thread; the whole-vector closures
// used by the CPU oracle call the same functions, so both sides share every expression.
template <int NS>
struct ModelHeat1dDae {
    static constexpr int N = NS, NP = 3;
    static constexpr bool HAS_MASS = true;
    static constexpr bool COMPONENTWISE = true;
    static constexpr int BAND_KL = 1, BAND_KU = 1;      // df/dy tridiagonal, M diagonal
    DSB_HD static double coef() { return 0.1 * (double)((NS - 1) * (NS - 1)); }
    template <class X>
    DSB_HD static double rhs_i(int i, const X& x, const double*, double) {
        if (i == 0 || i == NS - 1) return x[i];
        return coef() * (x[i - 1] - 2.0 * x[i] + x[i + 1]);
    }
    template <class X, class V>
    DSB_HD static double jac_mul_i(int i, const X&, const double*, double, const V& v) {
        if (i == 0 || i == NS - 1) return v[i];
        return coef() * (v[i - 1] - 2.0 * v[i] + v[i + 1]);
    }
    template <class X>
    DSB_HD static double mass_i(int i, const X& x, const double*, double, double beta, double yi) {
        if (i == 0 || i == NS - 1) return beta * yi;
        return x[i] + beta * yi;
    }
    DSB_HD static double init_i(int i, const double* p, double) {
        const double xi = (double)i / (double)(NS - 1);
        return (xi >= p[1] && xi <= p[2]) ? p[0] : 0.0;
    }
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) {
        for (int i = 0; i < N; ++i) y[i] = rhs_i(i, x, p, t);
    }
    DSB_HD static void jac_mul(const double* x, const double* p, double t, const double* v, double* y) {
        for (int i = 0; i < N; ++i) y[i] = jac_mul_i(i, x, p, t, v);
    }
    DSB_HD static void mass(const double* x, const double* p, double t, double beta, double* y) {
        for (int i = 0; i < N; ++i) y[i] = mass_i(i, x, p, t, beta, y[i]);
    }
    DSB_HD static void init(const double* p, double t, double* y) {
        for (int i = 0; i < N; ++i) y[i] = init_i(i, p, t);
    }
};

// The same equations with the boundary rows 0 = u - height / 4 while the initial profile is 0 at the boundary: the
// algebraic components START INCONSISTENT, so `new_and_consistent` has to move them (InitOp Newton with the backtracking
// line search, state.rs:84-162) before the first step -- the test problem of the initialisation kernels.
template <int NS>
struct ModelHeat1dDaeBc : ModelHeat1dDae<NS> {
    typedef ModelHeat1dDae<NS> Base;
    static constexpr int N = NS;
    template <class X>
    DSB_HD static double rhs_i(int i, const X& x, const double* p, double t) {
        if (i == 0 || i == NS - 1) return x[i] - 0.25 * p[0];
        return Base::rhs_i(i, x, p, t);
    }
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) {
        for (int i = 0; i < N; ++i) y[i] = rhs_i(i, x, p, t);
    }
};

// 2-D heat equation u_t = u_xx + u_yy on the unit square, MG x MG grid, 5-point differences on the interior points, the
// boundary rows algebraic (0 = u): the reference's heat2d test problem (test_models/heat2d.rs:101-196, the SUNDIALS
// idaHeat2D_klu example), whose statistics snapshot bdf.rs:2424-2446 is the only golden of the reference with n > 16.
// States only: the reference's output function (dx ||u||_2)^2 is evaluated by the tests from the states.
template <int MG>
struct ModelHeat2d {
    static constexpr int N = MG * MG, NP = 0;
    static constexpr bool HAS_MASS = true;
    static constexpr bool COMPONENTWISE = true;
    DSB_HD static bool boundary(int loc) { const int j = loc / MG, i = loc - j * MG; return j == 0 || j == MG - 1 || i == 0 || i == MG - 1; }
    DSB_HD static double coeff() { const double dx = 1.0 / ((double)MG - 1.0); return 1.0 / (dx * dx); }
    template <class X>
    DSB_HD static double rhs_i(int loc, const X& x, const double*, double) {
        if (boundary(loc)) return x[loc];
        return coeff() * (x[loc - 1] + x[loc + 1] + x[loc - MG] + x[loc + MG] - 4.0 * x[loc]);
    }
    template <class X, class V>
    DSB_HD static double jac_mul_i(int loc, const X&, const double*, double, const V& v) {
        if (boundary(loc)) return v[loc];
        return coeff() * (v[loc - 1] + v[loc + 1] + v[loc - MG] + v[loc + MG] - 4.0 * v[loc]);
    }
    template <class X>
    DSB_HD static double mass_i(int loc, const X& x, const double*, double, double beta, double yi) {
        if (boundary(loc)) return yi * beta;
        return x[loc] + beta * yi;
    }
    DSB_HD static double init_i(int loc, const double*, double) {
        if (boundary(loc)) return 0.0;
        const double dx = 1.0 / ((double)MG - 1.0);
        const int j = loc / MG, i = loc - j * MG;
        const double yfact = dx * (double)j, xfact = dx * (double)i;
        return 16.0 * xfact * (1.0 - xfact) * yfact * (1.0 - yfact);
    }
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) { for (int i = 0; i < N; ++i) y[i] = rhs_i(i, x, p, t); }
    DSB_HD static void jac_mul(const double* x, const double* p, double t, const double* v, double* y) { for (int i = 0; i < N; ++i) y[i] = jac_mul_i(i, x, p, t, v); }
    DSB_HD static void mass(const double* x, const double* p, double t, double beta, double* y) { for (int i = 0; i < N; ++i) y[i] = mass_i(i, x, p, t, beta, y[i]); }
    DSB_HD static void init(const double* p, double t, double* y) { for (int i = 0; i < N; ++i) y[i] = init_i(i, p, t); }
};

// Single-particle battery model (SPM) of the reference's battery example
// (examples/physics-based-battery-simulation/src/main.rs, model text book/src/primer/src/spm.ds), states only:
// u = [discharge capacity, throughput capacity, 20 negative-particle concentrations, 20 positive-particle
// concentrations]; p = [applied current I].  F is linear: two tridiagonal radial-diffusion operators plus a flux
// forcing on the surface node.  (The model's `out` / `stop` functions -- terminal voltage and its cut-offs --
// are not part of the implicit step loop and are not built.)  Products are summed in the order the model text
// lists the non-zeros of a row: super-diagonal, diagonal, sub-diagonal.
#include "dsb_spm_tables.inc"
struct dsb_spm_row { double sub, diag, sup; };
static const dsb_spm_row dsb_spm_neg_host[20] = { DSB_SPM_NEG_ROWS };
static const dsb_spm_row dsb_spm_pos_host[20] = { DSB_SPM_POS_ROWS };
static const dsb_spm_row dsb_spm99_neg_host[99] = { DSB_SPM99_NEG_ROWS };
static const dsb_spm_row dsb_spm99_pos_host[99] = { DSB_SPM99_POS_ROWS };
#if defined(__CUDACC__)
static __device__ const dsb_spm_row dsb_spm_neg_dev[20] = { DSB_SPM_NEG_ROWS };
static __device__ const dsb_spm_row dsb_spm_pos_dev[20] = { DSB_SPM_POS_ROWS };
static __device__ const dsb_spm_row dsb_spm99_neg_dev[99] = { DSB_SPM99_NEG_ROWS };
static __device__ const dsb_spm_row dsb_spm99_pos_dev[99] = { DSB_SPM99_POS_ROWS };
#endif
// the discretisation tables: 20 cells per particle (the reference's model text) or 99 (finite-volume refinement
// by the same formulas, tools/gen_spm_tables.py; not in the reference: BASELINE config 5 asks for n ~ 200)
struct SpmTables20 {
    static constexpr int NR = 20;
    DSB_HD static double neg_flux() { return DSB_SPM_NEG_FLUX; }
    DSB_HD static double pos_flux() { return DSB_SPM_POS_FLUX; }
    DSB_HD static dsb_spm_row neg(int k) {
#if defined(__CUDA_ARCH__)
        return dsb_spm_neg_dev[k];
#else
        return dsb_spm_neg_host[k];
#endif
    }
    DSB_HD static dsb_spm_row pos(int k) {
#if defined(__CUDA_ARCH__)
        return dsb_spm_pos_dev[k];
#else
        return dsb_spm_pos_host[k];
#endif
    }
};
struct SpmTables99 {
    static constexpr int NR = 99;
    DSB_HD static double neg_flux() { return DSB_SPM99_NEG_FLUX; }
    DSB_HD static double pos_flux() { return DSB_SPM99_POS_FLUX; }
    DSB_HD static dsb_spm_row neg(int k) {
#if defined(__CUDA_ARCH__)
        return dsb_spm99_neg_dev[k];
#else
        return dsb_spm99_neg_host[k];
#endif
    }
    DSB_HD static dsb_spm_row pos(int k) {
#if defined(__CUDA_ARCH__)
        return dsb_spm99_pos_dev[k];
#else
        return dsb_spm99_pos_host[k];
#endif
    }
};
// tuning of the warp-per-instance kernel for this model (measured, n = 200, 250 000 instances: 8 warps x chunk 2 (255 registers)
// 832 ms; 12 warps x chunk 1 (160 registers, no spills) 616 ms; 12 warps x chunk 2 spills 1 KB per lane: 976 ms)
#ifndef DSB_SPM_WBAND_WARPS
#define DSB_SPM_WBAND_WARPS 12
#endif
#ifndef DSB_SPM_WBAND_CHUNK
#define DSB_SPM_WBAND_CHUNK 1
#endif
template <class Tab>
struct ModelSpmT {
    static constexpr int NR = Tab::NR;
    static constexpr int N = 2 + 2 * NR, NP = 1;
    static constexpr bool HAS_MASS = false;
    static constexpr bool COMPONENTWISE = true;
    static constexpr int BAND_KL = 1, BAND_KU = 1;      // df/dy is tridiagonal (checked against the probed pattern at launch)
    static constexpr int WBAND_MAX_WARPS = DSB_SPM_WBAND_WARPS;   // warp-per-instance kernel (dsb_wband_bdf_kernel.cuh): warps per SM and
    static constexpr int WBAND_CHUNK = DSB_SPM_WBAND_CHUNK;       // components per chunk of its passes over the difference array
    template <class X>
    DSB_HD static double diffusion_i(int i, const X& x) {        // i in 2 .. N - 1
        const bool neg = i < 2 + NR;
        const int k = neg ? i - 2 : i - 2 - NR;
        const dsb_spm_row c = neg ? Tab::neg(k) : Tab::pos(k);
        double acc = 0.0;
        if (k < NR - 1) acc = c.sup * x[i + 1];
        acc = (k < NR - 1) ? (c.diag * x[i] + acc) : (c.diag * x[i]);
        if (k > 0) acc = c.sub * x[i - 1] + acc;
        return acc;
    }
    template <class X>
    DSB_HD static double rhs_i(int i, const X& x, const double* p, double) {
        if (i == 0) return 0.0002777777777777778 * p[0];
        if (i == 1) return 0.0002777777777777778 * dsb_abs(p[0]);
        const double flux = (i == 1 + NR) ? Tab::neg_flux() * (-520607810.21082705 * p[0])
                          : (i == 1 + 2 * NR) ? Tab::pos_flux() * (243644455.17866704 * p[0]) : 0.0;
        return diffusion_i(i, x) + flux;
    }
    template <class X, class V>
    DSB_HD static double jac_mul_i(int i, const X&, const double*, double, const V& v) {
        if (i < 2) return 0.0;
        return diffusion_i(i, v);
    }
    template <class X>
    DSB_HD static double mass_i(int i, const X& x, const double*, double, double beta, double yi) { return x[i] + beta * yi; }
    DSB_HD static double init_i(int i, const double*, double) {
        return i < 2 ? 0.0 : (i < 2 + NR ? 0.8000000000000016 : 0.6000000000000001);
    }
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) { for (int i = 0; i < N; ++i) y[i] = rhs_i(i, x, p, t); }
    DSB_HD static void jac_mul(const double* x, const double* p, double t, const double* v, double* y) { for (int i = 0; i < N; ++i) y[i] = jac_mul_i(i, x, p, t, v); }
    DSB_HD static void mass(const double* x, const double* p, double t, double beta, double* y) { for (int i = 0; i < N; ++i) y[i] = mass_i(i, x, p, t, beta, y[i]); }
    DSB_HD static void init(const double* p, double t, double* y) { for (int i = 0; i < N; ++i) y[i] = init_i(i, p, t); }
};
typedef ModelSpmT<SpmTables20> ModelSpm;
typedef ModelSpmT<SpmTables99> ModelSpm99;

// The battery model WITH its output function (spm.ds `out_i`: the terminal voltage) and its stop function (`stop_i`:
// the voltage leaves the window [3.105 V, 4.1 V]), i.e. with `OdeEquations::out` and `OdeEquations::root`: the integration ends at the first voltage cut-off exactly as in the reference's
// battery example (examples/physics-based-battery-simulation/src/main.rs: `RootFound(t, _) => finished`).
// voltage(x, I) restates `out_i` term by term, in the model text's order.  The surface concentrations are the
// two-point extrapolations of the model text (constant5 / 8 / 9 / 10: rows over the outermost two cells of a particle;
// products summed in the order the text lists them).  exp / tanh / arcsinh are the shared deterministic dsb_math.h
// versions, sqrt is correctly rounded: oracle and kernels agree bit for bit, the reference (libm) to ~1e-15.
template <class Tab>
struct ModelSpmStopT : ModelSpmT<Tab> {
    typedef ModelSpmT<Tab> Base;
    static constexpr int NR = Base::NR, N = Base::N;
    static constexpr int NROOTS = 2;
    template <class X>
    DSB_HD static double voltage(const X& x, const double* p) {
        const double cur = p[0];
        const double cn18 = x[2 + NR - 2], cn19 = x[2 + NR - 1];               // negative particle, outermost cells
        const double cp18 = x[2 + 2 * NR - 2], cp19 = x[2 + 2 * NR - 1];       // positive particle
        const double v2 = -25608.96286546366 * cp18 + 76826.88859639116 * cp19;
        const double v3 = -0.4999999999999983 * cp18 + 1.4999999999999982 * cp19;
        const double v4 = -12491.630996921805 * cn18 + 37474.892990765504 * cn19;
        const double v5 = -0.4999999999999983 * cn18 + 1.4999999999999984 * cn19;
        auto clamp = [](double v, double hi, double lo) { const double m = v < hi ? v : hi; return m > lo ? m : lo; };   // max(min(v, hi), lo)
        const double cps = clamp(v2, 51217.92521874824, 0.000512179257309275);
        const double xp = clamp(v3, 0.9999999999, 1e-10);
        const double cns = clamp(v4, 24983.261744011077, 0.000249832619938437);
        const double xn = clamp(v5, 0.9999999999, 1e-10);
        const double eta_p = 0.05138515824298745 * dsb_asinh((-2.3508116177110145 * cur)
                             / (2.0 * ((1.8973665961010275e-05 * dsb_sqrt(cps)) * dsb_sqrt(51217.9257309275 - cps))));
        double up = 2.16216 + 0.07645 * dsb_tanh(30.834 - 57.858397200000006 * xp);
        up = up + 2.1581 * dsb_tanh(52.294 - 53.412228 * xp);
        up = up - 0.14169 * dsb_tanh(11.0923 - 21.0852666 * xp);
        up = up + 0.2051 * dsb_tanh(1.4684 - 5.829105600000001 * xp);
        up = up + 0.2531 * dsb_tanh(4.291641337386018 - 8.069908814589667 * xp);
        up = up - 0.02167 * dsb_tanh(-87.5 + 177.0 * xp);
        up = up + 1e-06 * ((1.0 / xp) + (1.0 / (-1.0 + xp)));
        const double eta_n = 0.05138515824298745 * dsb_asinh((1.9590096814258458 * cur)
                             / (2.0 * ((0.0006324555320336759 * dsb_sqrt(cns)) * dsb_sqrt(24983.2619938437 - cns))));
        double un = 0.194 + 1.5 * dsb_exp(-120.0 * xn);
        un = un + 0.0351 * dsb_tanh(-3.44578313253012 + 12.048192771084336 * xn);
        un = un - 0.0045 * dsb_tanh(-7.1344537815126055 + 8.403361344537815 * xn);
        un = un - 0.035 * dsb_tanh(-18.466 + 20.0 * xn);
        un = un - 0.0147 * dsb_tanh(-14.705882352941176 + 29.41176470588235 * xn);
        un = un - 0.102 * dsb_tanh(-1.3661971830985917 + 7.042253521126761 * xn);
        un = un - 0.022 * dsb_tanh(-54.8780487804878 + 60.975609756097555 * xn);
        un = un - 0.011 * dsb_tanh(-5.486725663716814 + 44.24778761061947 * xn);
        un = un + 0.0155 * dsb_tanh(-3.6206896551724133 + 34.48275862068965 * xn);
        un = un + 1e-06 * ((1.0 / xn) + (1.0 / (-1.0 + xn)));
        return (eta_p + up) - (eta_n + un);
    }
    template <class X>
    DSB_HD static void root(const X& x, const double* p, double, double* g) {
        const double v = voltage(x, p);
        g[0] = -3.105 + v;
        g[1] = 4.1 - v;
    }
    // `out_i` of the model text: solve_dense returns the terminal voltage, one value per column (dense_write_out,
    // ode_solver/method.rs:822-848)
    static constexpr int NOUT = 1;
    template <class X>
    DSB_HD static void out(const X& x, const double* p, double, double* o) { o[0] = voltage(x, p); }
    // out and root read the outermost two cells of each particle only
    static constexpr int NDEP = 4;
    DSB_HD static int dep(int k) { return k == 0 ? 2 + NR - 2 : k == 1 ? 2 + NR - 1 : k == 2 ? 2 + 2 * NR - 2 : 2 + 2 * NR - 1; }
};
typedef ModelSpmStopT<SpmTables20> ModelSpmStop;
typedef ModelSpmStopT<SpmTables99> ModelSpm99Stop;

// The battery model cycled: the stop function's roots do not end the solve, the reset function (OdeEquations::reset,
// applied by solve_dense at every root, ode_solver/method.rs:783-797) puts the cell back into its initial, charged
// state and the discharge starts again.  Not a reference problem (the reference has no reset problem with n > 16): it
// exists to run resets through the banded lane kernels; pinned GPU-vs-oracle only.
template <class Tab>
struct ModelSpmCycleT : ModelSpmStopT<Tab> {
    typedef ModelSpmStopT<Tab> Base;
    static constexpr bool HAS_RESET = true;
    template <class X>
    DSB_HD static double reset_i(int i, const X&, const double* p, double t) { return Base::init_i(i, p, t); }
    DSB_HD static void reset(const double*, const double* p, double t, double* y) { for (int i = 0; i < Base::N; ++i) y[i] = Base::init_i(i, p, t); }
};
typedef ModelSpmCycleT<SpmTables20> ModelSpmCycle;

// id -> functor type
template <int ID> struct dsb_model_by_id;
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY> { typedef ModelExpDecay type; };
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY_ALGEBRAIC> { typedef ModelExpDecayAlgebraic type; };
template <> struct dsb_model_by_id<DSB_MODEL_ROBERTSON_DAE> { typedef ModelRobertsonDae type; };
template <> struct dsb_model_by_id<DSB_MODEL_ROBERTSON_ODE> { typedef ModelRobertsonOde<1> type; };
template <> struct dsb_model_by_id<DSB_MODEL_ROBERTSON_ODE_G3> { typedef ModelRobertsonOde<3> type; };
template <> struct dsb_model_by_id<DSB_MODEL_DYDT_Y2> { typedef ModelDydtY2<10> type; };
template <> struct dsb_model_by_id<DSB_MODEL_GAUSSIAN_DECAY> { typedef ModelGaussianDecay<10> type; };
template <> struct dsb_model_by_id<DSB_MODEL_VAN_DER_POL> { typedef ModelVanDerPol type; };
template <> struct dsb_model_by_id<DSB_MODEL_VAN_DER_POL_SCALED> { typedef ModelVanDerPolScaled type; };
template <> struct dsb_model_by_id<DSB_MODEL_HEAT1D_DAE_256> { typedef ModelHeat1dDae<256> type; };
template <> struct dsb_model_by_id<DSB_MODEL_HEAT1D_DAE_32> { typedef ModelHeat1dDae<32> type; };
template <> struct dsb_model_by_id<DSB_MODEL_SPM> { typedef ModelSpm type; };
template <> struct dsb_model_by_id<DSB_MODEL_SPM99> { typedef ModelSpm99 type; };
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY_ROOT> { typedef ModelExpDecayRoot type; };
template <> struct dsb_model_by_id<DSB_MODEL_SPM_STOP> { typedef ModelSpmStop type; };
template <> struct dsb_model_by_id<DSB_MODEL_SPM99_STOP> { typedef ModelSpm99Stop type; };
template <> struct dsb_model_by_id<DSB_MODEL_HEAT1D_DAE_32_BC> { typedef ModelHeat1dDaeBc<32> type; };
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY_RESET> { typedef ModelExpDecayReset type; };
template <> struct dsb_model_by_id<DSB_MODEL_HEAT2D_10> { typedef ModelHeat2d<10> type; };
template <> struct dsb_model_by_id<DSB_MODEL_BALL_BOUNCE> { typedef ModelBallBounce type; };
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY_TWO_ROOTS> { typedef ModelExpDecayTwoRoots type; };
template <> struct dsb_model_by_id<DSB_MODEL_SPM_CYCLE> { typedef ModelSpmCycle type; };
template <> struct dsb_model_by_id<DSB_MODEL_EXP_DECAY_ALGEBRAIC_RESET> { typedef ModelExpDecayAlgebraicReset type; };

// traits of an equation set: written component-wise (`*_i` functions), declares a band for df/dy
template <class M, class = void> struct dsb_is_componentwise : std::false_type {};
template <class M> struct dsb_is_componentwise<M, std::void_t<decltype(M::COMPONENTWISE)>> : std::bool_constant<M::COMPONENTWISE> {};
template <class M, class = void> struct dsb_declares_band : std::false_type {};
template <class M> struct dsb_declares_band<M, std::void_t<decltype(M::BAND_KL)>> : std::true_type {};

// Compile-time dispatch over the registry: calls f.template operator()<Model>() for `id`.
template <class F>
inline bool dsb_dispatch_model(int id, F&& f) {
    switch (id) {
        case DSB_MODEL_EXP_DECAY: f.template operator()<ModelExpDecay>(); return true;
        case DSB_MODEL_EXP_DECAY_ALGEBRAIC: f.template operator()<ModelExpDecayAlgebraic>(); return true;
        case DSB_MODEL_ROBERTSON_DAE: f.template operator()<ModelRobertsonDae>(); return true;
        case DSB_MODEL_ROBERTSON_ODE: f.template operator()<ModelRobertsonOde<1>>(); return true;
        case DSB_MODEL_ROBERTSON_ODE_G3: f.template operator()<ModelRobertsonOde<3>>(); return true;
        case DSB_MODEL_DYDT_Y2: f.template operator()<ModelDydtY2<10>>(); return true;
        case DSB_MODEL_GAUSSIAN_DECAY: f.template operator()<ModelGaussianDecay<10>>(); return true;
        case DSB_MODEL_VAN_DER_POL: f.template operator()<ModelVanDerPol>(); return true;
        case DSB_MODEL_VAN_DER_POL_SCALED: f.template operator()<ModelVanDerPolScaled>(); return true;
        case DSB_MODEL_HEAT1D_DAE_256: f.template operator()<ModelHeat1dDae<256>>(); return true;
        case DSB_MODEL_HEAT1D_DAE_32: f.template operator()<ModelHeat1dDae<32>>(); return true;
        case DSB_MODEL_SPM: f.template operator()<ModelSpm>(); return true;
        case DSB_MODEL_SPM99: f.template operator()<ModelSpm99>(); return true;
        case DSB_MODEL_EXP_DECAY_ROOT: f.template operator()<ModelExpDecayRoot>(); return true;
        case DSB_MODEL_SPM_STOP: f.template operator()<ModelSpmStop>(); return true;
        case DSB_MODEL_SPM99_STOP: f.template operator()<ModelSpm99Stop>(); return true;
        case DSB_MODEL_HEAT1D_DAE_32_BC: f.template operator()<ModelHeat1dDaeBc<32>>(); return true;
        case DSB_MODEL_EXP_DECAY_RESET: f.template operator()<ModelExpDecayReset>(); return true;
        case DSB_MODEL_HEAT2D_10: f.template operator()<ModelHeat2d<10>>(); return true;
        case DSB_MODEL_BALL_BOUNCE: f.template operator()<ModelBallBounce>(); return true;
        case DSB_MODEL_EXP_DECAY_TWO_ROOTS: f.template operator()<ModelExpDecayTwoRoots>(); return true;
        case DSB_MODEL_SPM_CYCLE: f.template operator()<ModelSpmCycle>(); return true;
        case DSB_MODEL_EXP_DECAY_ALGEBRAIC_RESET: f.template operator()<ModelExpDecayAlgebraicReset>(); return true;
        default: return false;
    }
}
