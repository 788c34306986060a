// dsb_math.h -- deterministic f64 math shared by the sm_100a kernels and the CPU oracle.
//
// Why this exists: the integer controller of diffsol's BDF/SDIRK loops branches on values of
// `powf` (reference: crates/diffsol-nl/src/convergence.rs:77,87,110;
// crates/diffsol/src/ode_solver/runge_kutta.rs:1313-1335; state.rs:1262-1263).  glibc `pow` and
// CUDA libdevice `pow` differ in the last ulp, which is enough to change step sizes and -- rarely --
// accepted-step counts.  Everything in this header uses only IEEE-754 +,-,*,/,fma,rint and integer
// bit manipulation, so a host build (gcc -ffp-contract=off) and a device build (nvcc --fmad=false)
// return bit-identical results.  Accuracy of dsb_pow is ~0.51 ulp (table-driven, double-double log,
// same construction idea as Tang 1989/1990); it agrees with a correctly rounded pow in ~99% of
// calls and is off by one ulp otherwise.
//
//   dsb_pow(x, y)   x >= 0 or NaN; replaces f64::powf on the hot path
//   dsb_powi(x, n)  square-and-multiply exactly as compiler-rt's __powidf2 (what f64::powi lowers to)
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define DSB_HD __host__ __device__ __forceinline__
#define DSB_HD_NOINLINE static __host__ __device__ __noinline__
#define DSB_SYMBOL extern "C" __host__ __device__      // a function of a DiffSL symbol table (dsb_diffsl_adapter.h)
#else
#define DSB_HD inline
#define DSB_HD_NOINLINE inline
#define DSB_SYMBOL extern "C"
#endif

#include "dsb_pow_tables.inc"

// number of root (event) functions of an equation set (csrc/dsb_models.h): M::NROOTS when declared, else 0.  Lives
// here because both the kernels and the oracle need it before the models are visible.
template <class M, class = void> struct dsb_model_nroots { static constexpr int value = 0; };
template <class M> struct dsb_model_nroots<M, decltype((void)M::NROOTS)> { static constexpr int value = M::NROOTS; };

// output function of an equation set (OdeEquations::out): M::NOUT outputs when declared, else solve_dense returns the states
template <class M, class = void> struct dsb_model_nout { static constexpr int value = M::N; static constexpr bool has_out = false; };
template <class M> struct dsb_model_nout<M, decltype((void)M::NOUT)> { static constexpr int value = M::NOUT; static constexpr bool has_out = true; };

// state components the out / root functions of an equation set read: M::NDEP of them, listed by M::dep(k), when the
// equations declare that (the lane kernels then interpolate only those for an output point or a root iteration); 0 = all
template <class M, class = void> struct dsb_model_ndep { static constexpr int value = 0; };
template <class M> struct dsb_model_ndep<M, decltype((void)M::NDEP)> { static constexpr int value = M::NDEP; };

// reset function of an equation set (OdeEquations::reset: the state map applied at a root): declared by M::HAS_RESET
template <class M, class = void> struct dsb_model_has_reset { static constexpr bool value = false; };
template <class M> struct dsb_model_has_reset<M, decltype((void)M::HAS_RESET)> { static constexpr bool value = M::HAS_RESET; };
// forward sensitivities (OdeEquationsImplicitSens, ode_equations/mod.rs): M::HAS_SENS, M::sens_mul(x, p, t, v, y) = f_p(x, p, t) v
// and M::init_sens(p, t, v, y) = (d y0 / d p) v
template <class M, class = void> struct dsb_model_has_sens { static constexpr bool value = false; };
template <class M> struct dsb_model_has_sens<M, decltype((void)M::HAS_SENS)> { static constexpr bool value = M::HAS_SENS; };
// DsbWithSens<M>: the equation set M with the sensitivity equations switched ON -- the kernels integrate one sensitivity
// vector per parameter beside the state (Bdf<.., SensEquations>, ode_solver/bdf.rs:934-989) only in this instantiation
template <class M, class = void> struct dsb_model_sens_on { static constexpr bool value = false; };
template <class M> struct dsb_model_sens_on<M, decltype((void)M::SENS_ON)> { static constexpr bool value = M::SENS_ON; };
template <class M> struct DsbWithSens : M { static constexpr bool SENS_ON = true; };
// DsbRagged<M>: the kernels in `solve(final_time)` form (OdeSolverMethod::solve, ode_solver/method.rs:227-258, 881-961) -- one
// result column per internal step, per instance as many as it takes, written at an offset the caller got from a counting
// pass.  A separate instantiation, so that the solve_dense kernels carry none of it.
template <class M, class = void> struct dsb_model_ragged_on { static constexpr bool value = false; };
template <class M> struct dsb_model_ragged_on<M, decltype((void)M::RAGGED_ON)> { static constexpr bool value = M::RAGGED_ON; };
template <class M> struct DsbRagged : M { static constexpr bool RAGGED_ON = true; };

struct dsb_log_row { double invc, logc_hi, logc_lo; };
struct dsb_exp_row { double hi, lo; };

static const dsb_log_row dsb_log_table_host[128] = { DSB_LOG_TABLE_ROWS };
static const dsb_exp_row dsb_exp_table_host[128] = { DSB_EXP_TABLE_ROWS };
#if defined(__CUDACC__)
static __device__ const dsb_log_row dsb_log_table_dev[128] = { DSB_LOG_TABLE_ROWS };
static __device__ const dsb_exp_row dsb_exp_table_dev[128] = { DSB_EXP_TABLE_ROWS };
#endif

DSB_HD uint64_t dsb_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
DSB_HD double dsb_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
DSB_HD double dsb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
DSB_HD double dsb_rint(double a) {
#if defined(__CUDA_ARCH__)
    return rint(a);
#else
    return __builtin_rint(a);
#endif
}
DSB_HD double dsb_sqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return __builtin_sqrt(a);
#endif
}
// IEEE division policies for the lane kernels.  On the device the inline expansion of `a / b` is ~21 SASS
// instructions per site.  The BDF lane kernel (30 sites; its loop body was instruction-cache bound) calls one shared
// copy instead: 1.14x faster on the Robertson sweep.  The SDIRK kernel fits the cache and keeps the inline form
// (the shared copy made it 1.18x slower: call overhead and lost instruction-level parallelism).
#if defined(__CUDACC__)
static __device__ __noinline__ double dsb_div_fn(double a, double b) { return a / b; }
struct DsbDivShared { static __device__ __forceinline__ double div(double a, double b) { return dsb_div_fn(a, b); } };
struct DsbDivInline { static __device__ __forceinline__ double div(double a, double b) { return a / b; } };
// N independent quotients in ONE call: the same IEEE divisions, but the N dependency chains (reciprocal seed, two Newton
// steps, quotient, residual correction: ~10 dependent operations each) interleave inside the routine instead of running
// one after the other through N calls.  The weighted norms divide N components at a time.
template <int N> struct DsbVecN { double v[N]; };
template <int N>
static __device__ __noinline__ DsbVecN<N> dsb_div_vec_fn(DsbVecN<N> a, DsbVecN<N> b) {
    DsbVecN<N> q;
#pragma unroll
    for (int i = 0; i < N; ++i) q.v[i] = a.v[i] / b.v[i];
    return q;
}
#endif
DSB_HD double dsb_abs(double a) { return dsb_from_bits(dsb_bits(a) & 0x7fffffffffffffffULL); }
DSB_HD bool dsb_isnan(double a) { return a != a; }

// ---- reciprocal-reuse division: a / d from a stored r = RN(1 / d), bit-identical to the IEEE quotient -------------
// The LU back-substitutions divide by the same diagonal in every Newton iteration between two factorisations, and the
// factorisation already forms RN(1 / diag) (nalgebra scales the sub-column by it).  Given that reciprocal the
// correctly rounded quotient takes 5 dependent operations instead of the ~10 (+ special-case test) of the full
// routine -- what matters in the band substitutions, whose recurrences are pure dependency chains:
//     q0 = RN(a r)                    relative error <= 2^-52 (two roundings)
//     e0 = RN(a - q0 d)  (fma)        the residual, itself with relative error <= 2^-53 if it needs 54 bits
//     q1 = RN(q0 + e0 r) (fma)        = RN(a/d + (a/d - q0) eps_r + ...): within 1/2 ulp + 2^-104 |a/d| of a/d, hence a
//                                     FAITHFUL rounding of a/d
//     e1 = a - q1 d      (fma)        exact: a faithful quotient's residual is representable
//     q  = RN(q1 + e1 r) (fma)        = RN(a / d) by Markstein's theorem (P. Markstein, IBM J. R&D 34 (1990); Muller et
//                                     al., Handbook of Floating-Point Arithmetic, section on division with an FMA):
//                                     a faithful q1 and r = RN(1/d) give the correctly rounded quotient in any binade
// The theorem needs every intermediate to stay normal: dsb_rcp() returns NaN unless |exponent(d)| <= 500, and
// dsb_div_rcp() falls back to the plain division unless |exponent(a)| <= 500 (zero, subnormal, infinite and NaN
// numerators take that path too, which also keeps the sign of a zero quotient) or when r is that NaN.  Checked against
// `/` on random and adversarial operands on the host (tests/test_oracle_golden.py::test_div_rcp_bit_exact) and on the
// device (tools/fp64_peak.cu).
DSB_HD double dsb_rcp(double d) {
    const uint32_t e = (uint32_t)(dsb_bits(d) >> 52) & 0x7ffu;
    const double r = 1.0 / d;
    return (e - (1023u - 500u) <= 1000u) ? r : dsb_from_bits(0x7ff8000000000000ULL);
}
// the fall-back is a real call so that the compiler cannot evaluate the full division speculatively and select
DSB_HD_NOINLINE double dsb_div_full(double a, double d) { return a / d; }
DSB_HD double dsb_div_rcp(double a, double d, double r) {
    const uint32_t e = (uint32_t)(dsb_bits(a) >> 52) & 0x7ffu;
    const double q0 = a * r;
    const double e0 = dsb_fma(-q0, d, a);
    const double q1 = dsb_fma(e0, r, q0);
    const double e1 = dsb_fma(-q1, d, a);
    const double q = dsb_fma(e1, r, q1);
    if (!(e - (1023u - 500u) <= 1000u && q == q)) return dsb_div_full(a, d);
    return q;
}
// r as dsb_rcp(d) would return it, from an already computed inv = 1.0 / d
DSB_HD double dsb_rcp_from(double d, double inv) {
    const uint32_t e = (uint32_t)(dsb_bits(d) >> 52) & 0x7ffu;
    return (e - (1023u - 500u) <= 1000u) ? inv : dsb_from_bits(0x7ff8000000000000ULL);
}
// The band back substitution of the warp-per-instance kernel (dsb_wband_bdf_kernel.cuh) splits the exponent budget
// unevenly -- pivots are O(1)-ish, right-hand sides span the far field of diffusion fronts: |exponent(d)| <= 100 and
// exponent(a) in [-920, 900] keep the quotient normal (>= -1020) and the residual a - q d exact (its last bit is
// 2^(exponent(a) - 104) >= 2^-1074).
DSB_HD double dsb_rcp_from_narrow(double d, double inv) {
    const uint32_t e = (uint32_t)(dsb_bits(d) >> 52) & 0x7ffu;
    return (e - (1023u - 100u) <= 200u) ? inv : dsb_from_bits(0x7ff8000000000000ULL);
}
// |q| with the sign bit of (a ^ d): one logic operation on the high word (the compiler would otherwise take |q| on the
// FP64 pipe, 8 cycles on the dependency chain)
DSB_HD double dsb_quotient_sign(double q, double a, double d) {
#if defined(__CUDA_ARCH__)
    const int sgn = (__double2hiint(a) ^ __double2hiint(d)) & (int)0x80000000;
    return __hiloint2double((__double2hiint(q) & 0x7fffffff) | sgn, __double2loint(q));
#else
    const uint64_t sgn = (dsb_bits(a) ^ dsb_bits(d)) & 0x8000000000000000ULL;
    return dsb_from_bits((dsb_bits(q) & 0x7fffffffffffffffULL) | sgn);
#endif
}
DSB_HD bool dsb_numerator_in_wide_range(double a) {         // exponent(a) in [-920, 900]: zero, subnormals, infinities, NaNs are out
    const uint32_t hi = (uint32_t)(dsb_bits(a) >> 32) & 0x7ff00000u;
    return hi - ((1023u - 920u) << 20) <= ((920u + 900u) << 20);
}

// Core of dsb_pow for a positive, finite, NORMAL x (bits ix) and finite y: exp(y * log(x)) with log(x) as
// a double-double (table of 128 sub-intervals, Tang-style) and a 128-entry 2^(j/128) table for exp.
DSB_HD double dsb_pow_core(uint64_t ix, int sub, double y) {
    const double inf = dsb_from_bits(0x7ff0000000000000ULL);
    // ---- log(x) = k ln2 + log(c_i) + log1p(r),  r = z/c_i - 1, as hi + lo ----
    const uint64_t OFF = 0x3FE6955500000000ULL;
    uint64_t tmp = ix - OFF;
    int i = (int)((tmp >> 45) & 127);
    int k = (int)((int64_t)tmp >> 52) - sub;
    uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
    double z = dsb_from_bits(iz);
    double kd = (double)k;
#if defined(__CUDA_ARCH__)
    const dsb_log_row row = dsb_log_table_dev[i];
#else
    const dsb_log_row row = dsb_log_table_host[i];
#endif
    // r = z*invc - 1 exactly as rhi + rlo
    double p_hi = z * row.invc;
    double p_lo = dsb_fma(z, row.invc, -p_hi);
    double q = p_hi - 1.0;                     // exact (Sterbenz)
    double r_hi = q + p_lo;
    double r_lo = (q - r_hi) + p_lo;           // fast two-sum: |q| >= |p_lo| or q == 0
    // -r^2/2 as a two-product
    double ar = -0.5 * r_hi;
    double s_hi = r_hi * ar;
    double s_lo = dsb_fma(r_hi, ar, -s_hi);
    // r^3 * (1/3 - r/4 + r^2/5 - r^3/6 + r^4/7 - r^5/8 + r^6/9)
    double r2 = r_hi * r_hi;
    double poly = 1.0 / 3.0 + r_hi * (-0.25 + r_hi * (0.2 + r_hi * (-1.0 / 6.0 + r_hi * (1.0 / 7.0
                  + r_hi * (-0.125 + r_hi * (1.0 / 9.0))))));
    double tail = (r2 * r_hi) * poly;
    // accumulate hi parts with two-sums, everything else into lo
    double a0 = kd * DSB_LN2_HI;               // exact: LN2_HI has 40 significant bits
    double t1 = a0 + row.logc_hi;
    double e1 = (a0 - t1) + row.logc_hi;       // |a0| >= |logc_hi| or a0 == 0
    double t2 = t1 + r_hi;
    double bb = t2 - t1;
    double e2 = (t1 - (t2 - bb)) + (r_hi - bb);
    double t3 = t2 + s_hi;
    bb = t3 - t2;
    double e3 = (t2 - (t3 - bb)) + (s_hi - bb);
    double lo = kd * DSB_LN2_LO + row.logc_lo;
    lo = lo + e1;
    lo = lo + e2;
    lo = lo + e3;
    lo = lo + r_lo;
    lo = lo + s_lo;
    lo = lo - r_hi * r_lo;                     // cross term of -(rhi+rlo)^2/2
    lo = lo + tail;
    double l_hi = t3 + lo;
    double l_lo = (t3 - l_hi) + lo;

    // ---- e = y * log(x) as hi + lo ----
    double e_hi = y * l_hi;
    double e_lo = dsb_fma(y, l_hi, -e_hi) + y * l_lo;

    // ---- exp(e_hi + e_lo) ----
    if (e_hi > 709.8) return inf;
    if (e_hi < -745.2) return 0.0;
    double zz = e_hi * DSB_EXP_INVLN2N;
    double kk = dsb_rint(zz);
    int64_t ki = (int64_t)kk;
    double rr = e_hi - kk * DSB_EXP_LN2N_HI;   // exact product (32-bit constant), exact difference
    rr = rr - kk * DSB_EXP_LN2N_LO;
    rr = rr + e_lo;
    int j = (int)(ki & 127);
    int64_t kq = ki >> 7;                      // floor division by 128
#if defined(__CUDA_ARCH__)
    const dsb_exp_row er = dsb_exp_table_dev[j];
#else
    const dsb_exp_row er = dsb_exp_table_host[j];
#endif
    double rr2 = rr * rr;
    double pe = rr + rr2 * (0.5 + rr * (1.0 / 6.0)) + (rr2 * rr2) * (1.0 / 24.0 + rr * (1.0 / 120.0 + rr * (1.0 / 720.0)));
    double val = er.hi + (er.lo + er.hi * pe);
    int64_t k1 = kq / 2;
    int64_t k2 = kq - k1;
    double sc1 = dsb_from_bits((uint64_t)(k1 + 1023) << 52);
    double sc2 = dsb_from_bits((uint64_t)(k2 + 1023) << 52);
    return (val * sc1) * sc2;
}

// Everything that is not (positive normal x, finite y): zeros, infinities, NaNs, negative and subnormal x.
DSB_HD_NOINLINE double dsb_pow_special(double x, double y) {
    const double inf = dsb_from_bits(0x7ff0000000000000ULL);
    const double nan = dsb_from_bits(0x7ff8000000000000ULL);
    if (y == 0.0) return 1.0;
    if (dsb_isnan(x) || dsb_isnan(y)) return nan;
    if (x < 0.0) return nan;
    if (x == 0.0) return y > 0.0 ? 0.0 : inf;
    if (x == inf) return y > 0.0 ? inf : 0.0;
    if (x == 1.0) return 1.0;
    if (y == inf) return x > 1.0 ? inf : 0.0;
    if (y == -inf) return x > 1.0 ? 0.0 : inf;
    uint64_t ix = dsb_bits(x);
    int sub = 0;
    if (ix < 0x0010000000000000ULL) {           // subnormal: scale by 2^52
        ix = dsb_bits(x * 4503599627370496.0);
        sub = 52;
    }
    return dsb_pow_core(ix, sub, y);
}

// x^y for x >= 0 (x < 0 returns NaN: the hot path never raises a negative base to a power).
DSB_HD_NOINLINE double dsb_pow(double x, double y) {
    if (y == 1.0) return x;                     // exact, and the most frequent call (Newton rate at the 2nd iteration)
    const uint64_t ix = dsb_bits(x);
    const uint64_t iy = dsb_bits(y) & 0x7fffffffffffffffULL;
    // fast path: x positive, finite and normal; y finite and non-zero
    const bool x_ok = (ix - 0x0010000000000000ULL) < (0x7ff0000000000000ULL - 0x0010000000000000ULL);
    const bool y_ok = (iy - 1ULL) < (0x7ff0000000000000ULL - 1ULL);
    if (!(x_ok && y_ok)) return dsb_pow_special(x, y);
    if (y == 0.5) return dsb_sqrt(x);           // correctly rounded
    return dsb_pow_core(ix, 0, y);
}

// f64::powi as lowered by LLVM on the reference's targets: compiler-rt / compiler_builtins
// __powidf2 (square-and-multiply, reciprocal at the end for negative exponents).
// Used at crates/diffsol-nl/src/convergence.rs:87 (`rate.pow(i32)`).
DSB_HD double dsb_powi(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

// ---- exp / log / tanh / asinh for model output and event functions (the battery model's terminal voltage) ----------
// Real calls on the device (one copy each): the voltage expression uses ~20 of them and is evaluated from several places
// of an integrator kernel; inlined everywhere the warp-per-instance kernel grew to 27 k SASS instructions and spent 85 %
// of its stall samples waiting for instruction fetch (profiles/r2_wband_spm_stop_inlined_math_hotspots.txt).
// Same tables and the same IEEE-only operation sequences as dsb_pow_core, so host and device agree bit for bit.  They
// are NOT correctly rounded (exp and log: below 1 ulp; tanh and asinh: absolute error of a few 1e-16, relative error
// that grows for |x| << 1), which is all an event threshold on a voltage needs; parity with the reference's libm
// values is a tolerance, parity between the oracle and the kernels is exact.
DSB_HD_NOINLINE double dsb_exp(double x) {
    const double inf = dsb_from_bits(0x7ff0000000000000ULL);
    if (x != x) return x;
    if (x > 709.8) return inf;
    if (x < -745.2) return 0.0;
    double zz = x * DSB_EXP_INVLN2N;
    double kk = dsb_rint(zz);
    int64_t ki = (int64_t)kk;
    double rr = x - kk * DSB_EXP_LN2N_HI;
    rr = rr - kk * DSB_EXP_LN2N_LO;
    int j = (int)(ki & 127);
    int64_t kq = ki >> 7;
#if defined(__CUDA_ARCH__)
    const dsb_exp_row er = dsb_exp_table_dev[j];
#else
    const dsb_exp_row er = dsb_exp_table_host[j];
#endif
    double rr2 = rr * rr;
    double pe = rr + rr2 * (0.5 + rr * (1.0 / 6.0)) + (rr2 * rr2) * (1.0 / 24.0 + rr * (1.0 / 120.0 + rr * (1.0 / 720.0)));
    double val = er.hi + (er.lo + er.hi * pe);
    int64_t k1 = kq / 2;
    int64_t k2 = kq - k1;
    double sc1 = dsb_from_bits((uint64_t)(k1 + 1023) << 52);
    double sc2 = dsb_from_bits((uint64_t)(k2 + 1023) << 52);
    return (val * sc1) * sc2;
}
// natural logarithm of a positive, finite, normal x (anything else: NaN for x < 0 or NaN, -inf for 0, x for +inf;
// subnormals are scaled first)
DSB_HD_NOINLINE double dsb_log(double x) {
    const double inf = dsb_from_bits(0x7ff0000000000000ULL);
    if (x != x || x < 0.0) return dsb_from_bits(0x7ff8000000000000ULL);
    if (x == 0.0) return -inf;
    if (x == inf) return x;
    uint64_t ix = dsb_bits(x);
    int sub = 0;
    if (ix < 0x0010000000000000ULL) { ix = dsb_bits(x * 4503599627370496.0); sub = 52; }
    const uint64_t OFF = 0x3FE6955500000000ULL;
    uint64_t tmp = ix - OFF;
    int i = (int)((tmp >> 45) & 127);
    int k = (int)((int64_t)tmp >> 52) - sub;
    uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
    double z = dsb_from_bits(iz);
    double kd = (double)k;
#if defined(__CUDA_ARCH__)
    const dsb_log_row row = dsb_log_table_dev[i];
#else
    const dsb_log_row row = dsb_log_table_host[i];
#endif
    double p_hi = z * row.invc;
    double p_lo = dsb_fma(z, row.invc, -p_hi);
    double q = p_hi - 1.0;
    double r_hi = q + p_lo;
    double r_lo = (q - r_hi) + p_lo;
    double ar = -0.5 * r_hi;
    double s_hi = r_hi * ar;
    double s_lo = dsb_fma(r_hi, ar, -s_hi);
    double r2 = r_hi * r_hi;
    double poly = 1.0 / 3.0 + r_hi * (-0.25 + r_hi * (0.2 + r_hi * (-1.0 / 6.0 + r_hi * (1.0 / 7.0
                  + r_hi * (-0.125 + r_hi * (1.0 / 9.0))))));
    double tail = (r2 * r_hi) * poly;
    double a0 = kd * DSB_LN2_HI;
    double t1 = a0 + row.logc_hi;
    double e1 = (a0 - t1) + row.logc_hi;
    double t2 = t1 + r_hi;
    double bb = t2 - t1;
    double e2 = (t1 - (t2 - bb)) + (r_hi - bb);
    double t3 = t2 + s_hi;
    bb = t3 - t2;
    double e3 = (t2 - (t3 - bb)) + (s_hi - bb);
    double lo = kd * DSB_LN2_LO + row.logc_lo;
    lo = lo + e1;
    lo = lo + e2;
    lo = lo + e3;
    lo = lo + r_lo;
    lo = lo + s_lo;
    lo = lo - r_hi * r_lo;
    lo = lo + tail;
    return t3 + lo;
}
DSB_HD_NOINLINE double dsb_tanh(double x) {
    if (x != x) return x;
    const double a = dsb_abs(x);
    if (a < 1e-8) return x;
    double r = 1.0;
    if (a < 20.0) r = 1.0 - 2.0 / (dsb_exp(2.0 * a) + 1.0);
    return x < 0.0 ? -r : r;
}
DSB_HD_NOINLINE double dsb_asinh(double x) {
    if (x != x) return x;
    const double a = dsb_abs(x);
    if (a < 1e-8) return x;
    const double r = (a > 1e100) ? dsb_log(a) + 0.6931471805599453 : dsb_log(a + dsb_sqrt(a * a + 1.0));
    return x < 0.0 ? -r : r;
}
