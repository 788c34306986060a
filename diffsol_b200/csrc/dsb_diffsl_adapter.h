// dsb_diffsl_adapter.h -- a DiffSL external module as an equation set of the batched kernels.
//
// The reference consumes compiled DiffSL models through a SYMBOL TABLE (crates/diffsol/src/ode_equations/diffsl.rs:
// 1072-1098 rhs / rhs_grad, 1221-1232 mass, 729-735 set_u0, 778-790 calc_stop, 972-984 calc_out, 405-418 set_inputs; the
// table itself: crates/diffsol-c/tests/external-dynamic-logistic/src/lib.rs:15 set_u0, :123 rhs, :142 rhs_grad, :233 mass,
// :303 calc_out, :407 calc_stop, :529 get_dims, :576 set_inputs).  A model that provides those functions with the same
// signatures -- as C source, every definition prefixed with DSB_SYMBOL so that it exists on the host AND on the device --
// is wrapped here into the functor interface of csrc/dsb_models.h (rhs / jac_mul / mass / init / root / out), compiled
// at run time into its own kernel-family instantiation (dsb_capi.cu: dsb_model_library_build) and into the CPU oracle
// (oracle/oracle.py: load_user_model), no enum entry and no rebuild of the library.
//
// The module's dimensions, which the reference reads through get_dims(), have to be compile-time constants for the
// kernels' register arrays; the source states them as
//     #define DSB_DIFFSL_STATES / _INPUTS / _OUTPUTS / _DATA / _STOP / _HAS_MASS / _HAS_SENS
// and get_dims() (optional) must agree: the loader calls it on the host build and checks.
//
// `data` -- the module's scratch block (inputs first, then whatever intermediates it keeps) -- is per solver in the
// reference; here every lane rebuilds it for each call (set_inputs, then set_u0 into a dummy state as
// DiffSl::set_params_and_model does, diffsl.rs:405-418), so that a call depends on nothing but its arguments.
// thread_id = 0, thread_dim = 1: the batch is parallel ACROSS instances.
#pragma once
#include "dsb_math.h"

// (DSB_SYMBOL comes from dsb_math.h: the model's source, which uses it, is included before this header)

#ifndef DSB_DIFFSL_OUTPUTS
#define DSB_DIFFSL_OUTPUTS 0
#endif
#ifndef DSB_DIFFSL_STOP
#define DSB_DIFFSL_STOP 0
#endif
#ifndef DSB_DIFFSL_HAS_MASS
#define DSB_DIFFSL_HAS_MASS 0
#endif
#ifndef DSB_DIFFSL_DATA
#define DSB_DIFFSL_DATA DSB_DIFFSL_INPUTS
#endif
#ifndef DSB_DIFFSL_HAS_SENS          // the module provides rhs_sgrad / set_u0_sgrad: forward sensitivities to its inputs
#define DSB_DIFFSL_HAS_SENS 0
#endif

// the symbol table (signatures of external-dynamic-logistic/src/lib.rs; u32 = unsigned)
DSB_SYMBOL void set_u0(double* u, double* data, unsigned thread_id, unsigned thread_dim);
DSB_SYMBOL void rhs(double t, const double* u, double* data, double* rr, unsigned thread_id, unsigned thread_dim);
DSB_SYMBOL void rhs_grad(double t, const double* u, const double* du, const double* data, double* ddata, const double* rr,
                         double* drr, unsigned thread_id, unsigned thread_dim);
DSB_SYMBOL void set_inputs(const double* inputs, double* data, unsigned model_index);
#if DSB_DIFFSL_HAS_MASS
DSB_SYMBOL void mass(double t, const double* v, double* data, double* mv, unsigned thread_id, unsigned thread_dim);
#endif
#if DSB_DIFFSL_OUTPUTS > 0
DSB_SYMBOL void calc_out(double t, const double* u, double* data, double* out, unsigned thread_id, unsigned thread_dim);
#endif
#if DSB_DIFFSL_STOP > 0
DSB_SYMBOL void calc_stop(double t, const double* u, double* data, double* root, unsigned thread_id, unsigned thread_dim);
#endif
#if DSB_DIFFSL_HAS_SENS
// forward-mode derivatives with respect to the inputs (external-dynamic-logistic/src/lib.rs:189 rhs_sgrad, :292 set_u0_sgrad)
DSB_SYMBOL void rhs_sgrad(double t, const double* u, const double* data, double* ddata, const double* rr, double* drr,
                          unsigned thread_id, unsigned thread_dim);
DSB_SYMBOL void set_u0_sgrad(const double* u, double* du, const double* data, double* ddata, unsigned thread_id, unsigned thread_dim);
#endif

struct DsbDiffslModel {
    static constexpr int N = DSB_DIFFSL_STATES, NP = DSB_DIFFSL_INPUTS;
    static constexpr int NDATA = DSB_DIFFSL_DATA > 0 ? DSB_DIFFSL_DATA : 1;
    static constexpr bool HAS_MASS = DSB_DIFFSL_HAS_MASS != 0;
    // DiffSl::set_params_and_model (diffsl.rs:405-418): the inputs into the data block, then set_u0 for the constants
    DSB_HD static void prepare(const double* p, double (&data)[NDATA]) {
#pragma unroll
        for (int k = 0; k < NDATA; ++k) data[k] = 0.0;
        double dummy[N];
        ::set_inputs(p, data, 0u);
        ::set_u0(dummy, data, 0u, 1u);
    }
    DSB_HD static void init(const double* p, double, double* y) {                  // DiffSlInit::call_inplace (diffsl.rs:729-735)
        double data[NDATA];
#pragma unroll
        for (int k = 0; k < NDATA; ++k) data[k] = 0.0;
        ::set_inputs(p, data, 0u);
        ::set_u0(y, data, 0u, 1u);
    }
    DSB_HD static void rhs(const double* x, const double* p, double t, double* y) { // DiffSlRhs::call_inplace (diffsl.rs:1073-1081)
        double data[NDATA];
        prepare(p, data);
        ::rhs(t, x, data, y, 0u, 1u);
    }
    // DiffSlRhs::jac_mul_inplace (diffsl.rs:1086-1098): ddata zeroed, tmp = the rhs scratch; the data block is the one rhs
    // left behind for this x (the reference calls rhs before any Jacobian product of a step)
    DSB_HD static void jac_mul(const double* x, const double* p, double t, const double* v, double* y) {
        double data[NDATA], ddata[NDATA], tmp[N];
        prepare(p, data);
        ::rhs(t, x, data, tmp, 0u, 1u);
#pragma unroll
        for (int k = 0; k < NDATA; ++k) ddata[k] = 0.0;
        ::rhs_grad(t, x, v, data, ddata, tmp, y, 0u, 1u);
    }
    // DiffSlMass::gemv_inplace (diffsl.rs:1221-1232): tmp = M x, then y = 1 * tmp + beta * y (nalgebra axpy)
    DSB_HD static void mass(const double* x, const double* p, double t, double beta, double* y) {
#if DSB_DIFFSL_HAS_MASS
        double data[NDATA], tmp[N];
        prepare(p, data);
        ::mass(t, x, data, tmp, 0u, 1u);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = 1.0 * tmp[i] + beta * y[i];
#else
        (void)p; (void)t;
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = x[i] + beta * y[i];
#endif
    }
#if DSB_DIFFSL_HAS_SENS
    static constexpr bool HAS_SENS = true;
    // DiffSlRhs::sens_mul_inplace (diffsl.rs:1152-1168): the direction v goes into the sensitivity data block through
    // set_inputs, then rhs_sgrad(t, x, data, sens_data, tmp, y) with tmp = the rhs scratch for this x
    DSB_HD static void sens_mul(const double* x, const double* p, double t, const double* v, double* y) {
        double data[NDATA], sdata[NDATA], tmp[N];
        prepare(p, data);
        ::rhs(t, x, data, tmp, 0u, 1u);
#pragma unroll
        for (int k = 0; k < NDATA; ++k) sdata[k] = 0.0;
        ::set_inputs(v, sdata, 0u);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = 0.0;
        ::rhs_sgrad(t, x, data, sdata, tmp, y, 0u, 1u);
    }
    // DiffSlInit::sens_mul_inplace (diffsl.rs:737-750): set_inputs(v -> sens_data), set_u0_sgrad(u0, y, data, sens_data) into
    // a zeroed y (ConstantOp::call allocates it with zeros)
    DSB_HD static void init_sens(const double* p, double, const double* v, double* y) {
        double data[NDATA], sdata[NDATA], u0[N];
#pragma unroll
        for (int k = 0; k < NDATA; ++k) { data[k] = 0.0; sdata[k] = 0.0; }
        ::set_inputs(p, data, 0u);
        ::set_u0(u0, data, 0u, 1u);
        ::set_inputs(v, sdata, 0u);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = 0.0;
        ::set_u0_sgrad(u0, y, data, sdata, 0u, 1u);
    }
#endif
#if DSB_DIFFSL_STOP > 0
    static constexpr int NROOTS = DSB_DIFFSL_STOP;
    template <class X>
    DSB_HD static void root(const X& x, const double* p, double t, double* g) {    // DiffSlRoot::call_inplace (diffsl.rs:778-790)
        double data[NDATA], xv[N];
        prepare(p, data);
#pragma unroll
        for (int i = 0; i < N; ++i) xv[i] = x[i];
        ::calc_stop(t, xv, data, g, 0u, 1u);
    }
#endif
#if DSB_DIFFSL_OUTPUTS > 0
    static constexpr int NOUT = DSB_DIFFSL_OUTPUTS;
    template <class X>
    DSB_HD static void out(const X& x, const double* p, double t, double* o) {     // DiffSlOut::call_inplace (diffsl.rs:972-984)
        double data[NDATA], xv[N];
        prepare(p, data);
#pragma unroll
        for (int i = 0; i < N; ++i) xv[i] = x[i];
        ::calc_out(t, xv, data, o, 0u, 1u);
    }
#endif
};
