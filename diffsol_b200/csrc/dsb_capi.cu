// dsb_capi.cu -- the extern "C" boundary declared in include/diffsol_b200.h.
//
// Host side of the batched integrator: problem description (what `OdeBuilder::build()` collects,
// crates/diffsol/src/ode_solver/builder.rs), device buffers of a batch, kernel launches, and the
// host<->device copies of the `*_host` entry points.  There is NO CPU fallback: every solve runs
// the sm_100a kernels; without a CUDA device the calls fail with DSB_ERR.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../include/diffsol_b200.h"
#include "dsb_args.h"
#include "dsb_host_setup.h"
#include "dsb_launch.h"
#include "dsb_lu_kernels.cuh"
#include "dsb_models.h"

// ---- thread-local last error (crates/diffsol-c/src/error_c.rs:12-46) -----------------------------------
static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define DSB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(DSB_ERR, std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

struct dsb_batch {
    dsb_problem prob;           // snapshot: the batch outlives edits of the problem
    int64_t B;
    int device;
    double* params = nullptr;   // [np][B]
    double* y0 = nullptr;
    double* dy0 = nullptr;
    double* h0 = nullptr;
    double* fin_t = nullptr;
    double* fin_h = nullptr;
    int32_t* fin_order = nullptr;
    int32_t* root_idx = nullptr;
    int32_t* ncols = nullptr;
    int32_t* stats = nullptr;   // [DSB_NSTATS][B]
    int32_t* status = nullptr;
    unsigned long long* work_counter = nullptr;
    double* t_eval = nullptr; int t_eval_cap = 0;
    double* ys_own = nullptr; size_t ys_own_bytes = 0;       // used by the *_host entry point
    double* sens_own = nullptr; size_t sens_own_bytes = 0;   // batch-major sensitivities of the *_sensitivities_host entry points
    int64_t* rag_off = nullptr;                              // [B + 1] column offsets of the last dsb_batch_solve_count
    int64_t rag_total = -1; int32_t rag_method = -1; double rag_final_time = 0.0;
    void* stage = nullptr; size_t stage_bytes = 0;           // instance-major staging for host copies
    cudaEvent_t ev0 = nullptr, ev_mid = nullptr, ev1 = nullptr;
    int last_launches = 0;
    bool have_timing = false;
    int sparsity_probe_jac_muls = 0;
    // the *_host entry points split big batches into chunks that run as child batches on their own streams, so that the
    // host -> device copy of one chunk's parameters and the device -> host copy of another's results overlap the kernels
    std::vector<dsb_batch*> chunks; std::vector<cudaStream_t> chunk_streams; std::vector<cudaEvent_t> chunk_done;
    DsbCoopState coop = {0, nullptr, 0, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, nullptr};
};

namespace {

// ---- equation sets loaded at run time (model plugins: dsb_inst.cu compiled for a user's source) -----------------------------
struct PluginModel {
    void* handle;
    dsb_launch_fn launch;
    void (*dims)(int*, int*, int*, int*);
    int (*coloring)(double, DsbProblemArgs*, int*, int32_t*, uint8_t*);
    std::string path;
};
std::vector<PluginModel>& plugins() { static std::vector<PluginModel> v; return v; }
const PluginModel* plugin_of(int model) {
    const int k = model - DSB_MODEL_PLUGIN_ID0;
    return (k >= 0 && k < (int)plugins().size()) ? &plugins()[(size_t)k] : nullptr;
}
// fill_problem_args (dsb_host_setup.h) asks here for the sparsity pattern and colouring of a plugin model
int plugin_coloring_hook(const dsb_problem& pr, DsbProblemArgs* pa, int* probes, std::vector<int32_t>* color_full, std::vector<uint8_t>* nz_full) {
    const PluginModel* pm = plugin_of(pr.model);
    if (!pm) return DSB_BAD_ARG;
    color_full->assign((size_t)pr.n, 0); nz_full->assign((size_t)pr.n * pr.n, 0);
    pm->coloring(pr.t0, pa, probes, color_full->data(), nz_full->data());
    return DSB_OK;
}

struct DimsOf {
    int *n, *np, *hm, *nout;
    template <class M> void operator()() { *n = M::N; *np = M::NP; *hm = M::HAS_MASS ? 1 : 0; *nout = dsb_model_nout<M>::value; }
};

using namespace dsb_host;

const dsb_launch_fn g_launch_table[DSB_MODEL_COUNT] = {
    dsb_launch_model_0, dsb_launch_model_1, dsb_launch_model_2, dsb_launch_model_3,
    dsb_launch_model_4, dsb_launch_model_5, dsb_launch_model_6, dsb_launch_model_7,
    dsb_launch_model_8, dsb_launch_model_9, dsb_launch_model_10, dsb_launch_model_11,
    dsb_launch_model_12, dsb_launch_model_13, dsb_launch_model_14, dsb_launch_model_15, dsb_launch_model_16, dsb_launch_model_17, dsb_launch_model_18, dsb_launch_model_19, dsb_launch_model_20, dsb_launch_model_21, dsb_launch_model_22,
};

// instance-major <-> batch-major re-layout on the device (the host-facing layouts follow the
// reference: parameters concatenated per instance, each instance's solve_dense block column-major)
__global__ void dsb_to_batch_major_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t B, int m) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*m, dst index
    if (idx >= B * m) return;
    const int64_t j = idx / B, b = idx - j * B;
    dst[idx] = src[b * m + j];
}
__global__ void dsb_to_instance_major_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t B, int m) {
    __shared__ double tile[32][33];
    // src [m][B] -> dst [B][m], 32x32 tiles through shared memory so both sides are coalesced
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += blockDim.y) {
        const int j = j0 + r; const int64_t b = b0 + tx;
        if (j < m && b < B) tile[r][tx] = src[(int64_t)j * B + b];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += blockDim.y) {
        const int64_t b = b0 + r; const int j = j0 + tx;
        if (j < m && b < B) dst[b * m + j] = tile[tx][r];
    }
}
__global__ void dsb_im_to_batch_major_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t B, int m) {
    __shared__ double tile[32][33];
    // src [B][m] -> dst [m][B], 32x32 tiles through shared memory so both sides are coalesced
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += blockDim.y) {
        const int64_t b = b0 + r; const int j = j0 + tx;
        if (j < m && b < B) tile[r][tx] = src[b * m + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += blockDim.y) {
        const int j = j0 + r; const int64_t b = b0 + tx;
        if (j < m && b < B) dst[(int64_t)j * B + b] = tile[tx][r];
    }
}
__global__ void dsb_stats_to_host_layout_kernel(const int32_t* __restrict__ src, int64_t* __restrict__ dst, int64_t B,
                                                int probe_jac_muls) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*DSB_NSTATS, dst index
    if (idx >= B * DSB_NSTATS) return;
    const int64_t b = idx / DSB_NSTATS; const int s = (int)(idx - b * DSB_NSTATS);
    int64_t v = src[(int64_t)s * B + b];
    if (s == DSB_STAT_RHS_JAC_MULS) v += probe_jac_muls;
    dst[idx] = v;
}
__global__ void dsb_fill_i32_kernel(int32_t* __restrict__ dst, int64_t B, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) dst[i] = v;
}
__global__ void dsb_sum_stat_kernel(const int32_t* __restrict__ src, int64_t B, unsigned long long* out) {
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
        acc += (unsigned long long)src[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

int ensure_stage(dsb_batch* b, size_t bytes) {
    if (b->stage_bytes >= bytes) return DSB_OK;
    if (b->stage) cudaFree(b->stage);
    b->stage = nullptr; b->stage_bytes = 0;
    DSB_CUDA(cudaMalloc(&b->stage, bytes));
    b->stage_bytes = bytes;
    return DSB_OK;
}

}  // namespace

extern "C" {

void dsb_options_default(dsb_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->max_nonlinear_solver_iterations = 10;
    o->max_error_test_failures = 40;
    o->max_nonlinear_solver_failures = 50;
    o->update_jacobian_after_steps = 20;
    o->update_rhs_jacobian_after_steps = 50;
    o->ic_max_linesearch_iterations = 10;
    o->ic_max_newton_iterations = 10;
    o->ic_max_linear_solver_setups = 4;
    o->ic_use_linesearch = 1;
    o->nonlinear_solver_tolerance = 0.2;
    o->min_timestep = 1e-13;
    o->max_timestep_growth = 2.0;
    o->min_timestep_growth = 2.0;
    o->max_timestep_shrink = 0.9;
    o->min_timestep_shrink = 0.5;
    o->threshold_to_update_jacobian = 0.3;
    o->threshold_to_update_rhs_jacobian = 0.2;
    o->pi_control_proportional = 0.0;
    o->pi_control_integral = 0.5;
    o->ic_step_reduction_factor = 0.5;
    o->ic_armijo_constant = 1e-4;
}

const char* dsb_last_error(void) { return g_last_error.c_str(); }
const char* dsb_version(void) { return "diffsol_b200 0.1 (sm_100a)"; }

int dsb_device_count(int* count) {
    if (!count) return fail(DSB_BAD_ARG, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess || c == 0) {
        *count = 0;
        return fail(DSB_ERR, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    *count = c;
    return DSB_OK;
}

int dsb_problem_new(int model, dsb_problem** out) {
    if (!out) return fail(DSB_BAD_ARG, "out is NULL");
    int n = 0, np = 0, hm = 0, nout = 0;
    DimsOf f{&n, &np, &hm, &nout};
    if (const PluginModel* pm = plugin_of(model)) pm->dims(&n, &np, &hm, &nout);
    else if (!dsb_dispatch_model(model, f)) return fail(DSB_BAD_ARG, "unknown model id");
    dsb_problem* p = new (std::nothrow) dsb_problem();
    if (!p) return fail(DSB_ERR, "out of memory");
    p->model = model; p->n = n; p->np = np; p->has_mass = hm; p->nout = nout;
    p->rtol = 1e-6; p->atol.assign(1, 1e-6); p->t0 = 0.0; p->h0 = 1.0; p->use_coloring = 0;   // builder.rs:112-140
    dsb_options_default(&p->opt);
    *out = p;
    return DSB_OK;
}
int dsb_problem_free(dsb_problem* p) { delete p; return DSB_OK; }
int dsb_problem_dims(const dsb_problem* p, int32_t* nstates, int32_t* nparams, int32_t* has_mass) {
    if (!p) return fail(DSB_BAD_ARG, "problem is NULL");
    if (nstates) *nstates = p->n;
    if (nparams) *nparams = p->np;
    if (has_mass) *has_mass = p->has_mass;
    return DSB_OK;
}
int dsb_problem_nout(const dsb_problem* p, int32_t* nout) {
    if (!p || !nout) return fail(DSB_BAD_ARG, "NULL argument");
    *nout = p->nout;
    return DSB_OK;
}
int dsb_problem_set_rtol(dsb_problem* p, double rtol) {
    if (!p || !(rtol > 0.0)) return fail(DSB_BAD_ARG, "rtol must be positive");
    p->rtol = rtol; return DSB_OK;
}
int dsb_problem_set_atol(dsb_problem* p, const double* atol, int32_t n) {
    if (!p || !atol || (n != 1 && n != p->n)) return fail(DSB_BAD_ARG, "atol must have 1 or nstates entries");
    p->atol.assign(atol, atol + n); return DSB_OK;
}
int dsb_problem_set_t0(dsb_problem* p, double t0) { if (!p) return fail(DSB_BAD_ARG, "problem is NULL"); p->t0 = t0; return DSB_OK; }
int dsb_problem_set_h0(dsb_problem* p, double h0) {
    if (!p || h0 == 0.0) return fail(DSB_BAD_ARG, "h0 must be non-zero");
    p->h0 = h0; return DSB_OK;
}
int dsb_problem_set_use_coloring(dsb_problem* p, int32_t use_coloring) {
    if (!p) return fail(DSB_BAD_ARG, "problem is NULL");
    p->use_coloring = use_coloring ? 1 : 0; return DSB_OK;
}
int dsb_problem_set_options(dsb_problem* p, const dsb_options* opt) {
    if (!p || !opt) return fail(DSB_BAD_ARG, "NULL argument");
    if (opt->max_nonlinear_solver_iterations < 1) return fail(DSB_BAD_ARG, "max_nonlinear_solver_iterations < 1");
    p->opt = *opt; return DSB_OK;
}
int dsb_problem_set_sensitivities(dsb_problem* p, int32_t enable, double sens_rtol, const double* sens_atol, int32_t natol) {
    if (!p) return fail(DSB_BAD_ARG, "problem is NULL");
    if (!enable) { p->sens = 0; p->sens_atol.clear(); p->sens_rtol = 0.0; return DSB_OK; }
    if (natol != 0 && natol != 1 && natol != p->n) return fail(DSB_BAD_ARG, "sens_atol must have 0, 1 or nstates entries");
    if (natol > 0 && (!sens_atol || !(sens_rtol >= 0.0))) return fail(DSB_BAD_ARG, "sens_atol is NULL or sens_rtol is negative");
    p->sens = 1;
    p->sens_rtol = sens_rtol;
    p->sens_atol.assign(sens_atol, sens_atol + (natol > 0 ? natol : 0));
    return DSB_OK;
}
int dsb_problem_get_options(const dsb_problem* p, dsb_options* opt) {
    if (!p || !opt) return fail(DSB_BAD_ARG, "NULL argument");
    *opt = p->opt; return DSB_OK;
}

// ---- user equation sets ("DiffSL modules drop in") -------------------------------------------------------
// The source is compiled by nvcc for sm_100a into ONE instantiation of the kernel families (csrc/dsb_inst.cu with the
// user's text in place of a built-in functor) and linked into a shared object that this library loads; the kernels that
// integrate a user model are therefore the same hand-written kernels, specialised by the compiler for its size and
// functions -- no interpreter, no enum entry, no rebuild of this library.
int dsb_model_library_build(const char* source_path, int32_t kind, const char* struct_name, const char* csrc_dir,
                            const char* out_path) {
    if (!source_path || !csrc_dir || !out_path) return fail(DSB_BAD_ARG, "NULL argument");
    if (kind != DSB_MODEL_SOURCE_FUNCTOR && kind != DSB_MODEL_SOURCE_DIFFSL) return fail(DSB_BAD_ARG, "unknown source kind");
    if (kind == DSB_MODEL_SOURCE_FUNCTOR && (!struct_name || !*struct_name)) return fail(DSB_BAD_ARG, "struct_name is required for a functor source");
    for (const char* q : {source_path, csrc_dir, out_path, struct_name ? struct_name : ""})
        for (const char* c = q; *c; ++c)
            if (*c == '"' || *c == '\'' || *c == '`' || *c == '$' || *c == ';' || *c == '\n') return fail(DSB_BAD_ARG, "unsupported character in a path or name");
    const char* nvcc = getenv("NVCC");
    std::string cmd = std::string(nvcc && *nvcc ? nvcc : "nvcc") +
        " -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math"
        " -diag-suppress 128 -shared -I'" + csrc_dir + "' -DDSB_USER_MODEL_SOURCE='\"" + source_path + "\"'";
    if (kind == DSB_MODEL_SOURCE_DIFFSL) cmd += " -DDSB_USER_DIFFSL";
    else cmd += std::string(" -DDSB_USER_MODEL=") + struct_name;
    cmd += std::string(" '") + csrc_dir + "/dsb_inst.cu' -o '" + out_path + "' 2>&1";
    FILE* pipe = popen(cmd.c_str(), "r");
    if (!pipe) return fail(DSB_ERR, "cannot run nvcc");
    std::string log; char buf[512];
    while (fgets(buf, sizeof(buf), pipe)) log += buf;
    const int rc = pclose(pipe);
    if (rc != 0) return fail(DSB_ERR, "nvcc failed on the model source:\n" + log.substr(0, 4000));
    return DSB_OK;
}

int dsb_model_library_load(const char* library_path, int32_t* model_out) {
    if (!library_path || !model_out) return fail(DSB_BAD_ARG, "NULL argument");
    for (size_t k = 0; k < plugins().size(); ++k)
        if (plugins()[k].path == library_path) { *model_out = DSB_MODEL_PLUGIN_ID0 + (int)k; return DSB_OK; }
    void* h = dlopen(library_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(DSB_ERR, std::string("dlopen: ") + dlerror());
    PluginModel pm;
    pm.handle = h; pm.path = library_path;
    int (*abi)(void) = (int (*)(void))dlsym(h, "dsb_plugin_abi");
    pm.launch = (dsb_launch_fn)dlsym(h, "dsb_plugin_launch");
    pm.dims = (void (*)(int*, int*, int*, int*))dlsym(h, "dsb_plugin_dims");
    pm.coloring = (int (*)(double, DsbProblemArgs*, int*, int32_t*, uint8_t*))dlsym(h, "dsb_plugin_coloring");
    if (!abi || !pm.launch || !pm.dims || !pm.coloring) { dlclose(h); return fail(DSB_ERR, "not a diffsol_b200 model library (symbols missing)"); }
    if (abi() != 2) { dlclose(h); return fail(DSB_ERR, "model library built against another version of csrc/ (rebuild it)"); }
    plugins().push_back(pm);
    dsb_host::plugin_coloring() = &plugin_coloring_hook;
    *model_out = DSB_MODEL_PLUGIN_ID0 + (int)plugins().size() - 1;
    return DSB_OK;
}

int dsb_batch_new(const dsb_problem* p, int64_t nbatch, int32_t device, dsb_batch** out) {
    if (!p || !out || nbatch < 1) return fail(DSB_BAD_ARG, "bad argument to dsb_batch_new");
    if (nbatch > 0x7fffffffll * 64) return fail(DSB_BAD_ARG, "nbatch too large");
    int count = 0;
    if (dsb_device_count(&count) != DSB_OK) return DSB_ERR;
    if (device < 0 || device >= count) return fail(DSB_BAD_ARG, "no such CUDA device");
    DSB_CUDA(cudaSetDevice(device));
    dsb_batch* b = new (std::nothrow) dsb_batch();
    if (!b) return fail(DSB_ERR, "out of memory");
    b->prob = *p; b->B = nbatch; b->device = device;
    const size_t B = (size_t)nbatch, n = (size_t)p->n, np = (size_t)(p->np > 0 ? p->np : 1);
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
    alloc((void**)&b->params, np * B * 8);
    alloc((void**)&b->y0, n * B * 8);
    alloc((void**)&b->dy0, n * B * 8);
    alloc((void**)&b->h0, B * 8);
    alloc((void**)&b->fin_t, B * 8);
    alloc((void**)&b->fin_h, B * 8);
    alloc((void**)&b->fin_order, B * 4);
    alloc((void**)&b->root_idx, B * 4);
    alloc((void**)&b->ncols, B * 4);
    alloc((void**)&b->stats, (size_t)DSB_NSTATS * B * 4);
    alloc((void**)&b->status, B * 4);
    alloc((void**)&b->work_counter, 256);     // word 0: the work counter; words 1..31: diagnostics (DSB_LANE_PROFILE builds)
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_mid);
    if (e != cudaSuccess) {
        dsb_batch_free(b);
        return fail(DSB_ERR, std::string("dsb_batch_new: ") + cudaGetErrorString(e));
    }
    cudaMemset(b->params, 0, np * B * 8);
    cudaMemset(b->stats, 0, (size_t)DSB_NSTATS * B * 4);
    cudaMemset(b->status, 0, B * 4);
    cudaMemset(b->fin_t, 0, B * 8); cudaMemset(b->fin_h, 0, B * 8); cudaMemset(b->fin_order, 0, B * 4);
    *out = b;
    return DSB_OK;
}

int dsb_batch_free(dsb_batch* b) {
    if (!b) return DSB_OK;
    cudaSetDevice(b->device);
    for (dsb_batch* c : b->chunks) dsb_batch_free(c);
    for (cudaStream_t st : b->chunk_streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : b->chunk_done) cudaEventDestroy(ev);
    cudaFree(b->params); cudaFree(b->y0); cudaFree(b->dy0); cudaFree(b->h0);
    cudaFree(b->fin_t); cudaFree(b->fin_h); cudaFree(b->fin_order); cudaFree(b->root_idx); cudaFree(b->ncols);
    cudaFree(b->stats); cudaFree(b->status); cudaFree(b->work_counter); cudaFree(b->coop.ws_mem); cudaFree(b->coop.atol_dev); cudaFree(b->coop.color_dev); cudaFree(b->coop.wb_mem); cudaFree(b->coop.ys_im_own); cudaFree(b->t_eval); cudaFree(b->ys_own); cudaFree(b->sens_own); cudaFree(b->rag_off); cudaFree(b->stage);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->ev_mid) cudaEventDestroy(b->ev_mid);
    delete b;
    return DSB_OK;
}
int64_t dsb_batch_size(const dsb_batch* b) { return b ? b->B : 0; }

int dsb_batch_set_params_device(dsb_batch* b, const double* params_dev, int64_t nbatch, int32_t nparams, void* stream) {
    if (!b || nbatch != b->B || nparams != b->prob.np) return fail(DSB_BAD_ARG, "parameter shape mismatch");
    if (nparams == 0) return DSB_OK;
    if (!params_dev) return fail(DSB_BAD_ARG, "params is NULL");
    DSB_CUDA(cudaSetDevice(b->device));
    const int64_t total = nbatch * nparams;
    dsb_to_batch_major_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params_dev, b->params, nbatch, nparams);
    DSB_CUDA(cudaGetLastError());
    return DSB_OK;
}

int dsb_batch_set_params_host(dsb_batch* b, const double* params, int64_t nbatch, int32_t nparams) {
    if (!b || nbatch != b->B || nparams != b->prob.np) return fail(DSB_BAD_ARG, "parameter shape mismatch");
    if (nparams == 0) return DSB_OK;
    if (!params) return fail(DSB_BAD_ARG, "params is NULL");
    DSB_CUDA(cudaSetDevice(b->device));
    const size_t bytes = (size_t)nbatch * nparams * 8;
    if (ensure_stage(b, bytes) != DSB_OK) return DSB_ERR;
    DSB_CUDA(cudaMemcpyAsync(b->stage, params, bytes, cudaMemcpyHostToDevice, 0));
    int rc = dsb_batch_set_params_device(b, (const double*)b->stage, nbatch, nparams, nullptr);
    if (rc != DSB_OK) return rc;
    DSB_CUDA(cudaStreamSynchronize(0));
    return DSB_OK;
}

// ys_im: instance-major buffer offered to kernels that write that layout (NULL: they use their own and the result is
// re-laid out into ys_dev); *wrote_im (may be NULL) reports that the result is in ys_im and ys_dev was NOT written
// what the sensitivity and solve(final_time) entry points add to a solve
struct SolveExtras {
    double* sens_dev = nullptr;                     // solve_dense_sensitivities: [nt][np][n][B]
    int ragged = 0;                                 // solve(final_time): 1 = counting pass, 2 = writing pass
    const int64_t* rag_off = nullptr; double* rag_ts = nullptr; double* rag_ys = nullptr;
};
static int solve_impl(dsb_batch* b, int32_t method, const double* t_eval, int32_t nt, double* ys_dev, void* stream_,
                      int free_running, double* ys_im = nullptr, int* wrote_im = nullptr, const SolveExtras* ex = nullptr) {
    double* const sens_dev = ex ? ex->sens_dev : nullptr;
    if (!b || !t_eval || nt < 1 || !ys_dev) return fail(DSB_BAD_ARG, "bad argument to dsb_batch_solve_dense");
    if (sens_dev && !b->prob.sens) return fail(DSB_BAD_ARG, "the problem has no sensitivities enabled (dsb_problem_set_sensitivities)");
    if (!sens_dev && b->prob.sens) return fail(DSB_BAD_ARG, "a problem with sensitivities is solved through dsb_batch_solve_dense_sensitivities");
    if (method != DSB_METHOD_BDF && method != DSB_METHOD_TR_BDF2 && method != DSB_METHOD_ESDIRK34)
        return fail(DSB_BAD_ARG, "unknown method");
    // solve_dense walks forward through t_eval (method.rs:761-764); the free-running loop steps while |t| < |t_k|
    // (ode_solver/mod.rs:132-141), which also serves backward integration (h0 < 0: negative_exponential_decay_problem)
    for (int k = 1; k < nt; ++k) {
        if (!free_running && !(t_eval[k] >= t_eval[k - 1])) return fail(DSB_BAD_ARG, "t_eval must be increasing");
        if (free_running && !(std::fabs(t_eval[k]) >= std::fabs(t_eval[k - 1]))) return fail(DSB_BAD_ARG, "|t_points| must be increasing");
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    DSB_CUDA(cudaSetDevice(b->device));
    if (b->t_eval_cap < nt) {
        cudaFree(b->t_eval); b->t_eval = nullptr; b->t_eval_cap = 0;
        DSB_CUDA(cudaMalloc((void**)&b->t_eval, (size_t)nt * 8));
        b->t_eval_cap = nt;
    }
    DSB_CUDA(cudaMemcpyAsync(b->t_eval, t_eval, (size_t)nt * 8, cudaMemcpyHostToDevice, stream));
    DsbProblemArgs pa;
    int probes = 0;
    std::vector<int32_t> color_full; std::vector<uint8_t> nz_full;
    if (fill_problem_args(b->prob, b->B, nt, &pa, &probes, &color_full, &nz_full) != DSB_OK) return fail(DSB_BAD_ARG, "unknown model id");
    b->sparsity_probe_jac_muls = probes;
    pa.free_running = free_running;
    build_tableau(method, &pa.rk);
    pa.quorum = DSB_DEFAULT_QUORUM;
    pa.newton_passes = DSB_DEFAULT_NEWTON_PASSES;
    if (const char* q = getenv("DSB_NEWTON_PASSES")) { int v = atoi(q); if (v >= 1 && v <= 16) pa.newton_passes = v; }   // tuning knob
    if (const char* q = getenv("DSB_COOP_DENSE_ONLY")) pa.coop_dense_only = atoi(q);
    if (const char* q = getenv("DSB_WBAND_FORCE_REDO")) pa.reserved1 = atoi(q);      // test hook (dsb_wband_bdf_kernel.cuh)
    if (const char* q = getenv("DSB_QUORUM")) { int v = atoi(q); if (v >= 1 && v <= 33) pa.quorum = v; }   // tuning knob
    DsbBatchBuffers bb;
    bb.params = b->params; bb.t_eval = b->t_eval; bb.y0 = b->y0; bb.dy0 = b->dy0; bb.h0 = b->h0;
    bb.ys = ys_dev; bb.stats = b->stats; bb.status = b->status;
    bb.fin_t = b->fin_t; bb.fin_h = b->fin_h; bb.fin_order = b->fin_order;
    bb.root_idx = b->root_idx; bb.ncols = b->ncols;
    bb.ss = sens_dev;
    bb.rag_off = ex ? ex->rag_off : nullptr; bb.rag_ts = ex ? ex->rag_ts : nullptr; bb.rag_ys = ex ? ex->rag_ys : nullptr;
    pa.ragged = ex ? ex->ragged : 0;
    if (sens_dev) DSB_CUDA(cudaMemsetAsync(sens_dev, 0xFF, (size_t)nt * b->prob.np * b->prob.n * b->B * 8, stream));
    // kernels without root finding leave these alone: no root, every column
    DSB_CUDA(cudaMemsetAsync(b->root_idx, 0xFF, (size_t)b->B * 4, stream));
    dsb_fill_i32_kernel<<<(unsigned)((b->B + 255) / 256), 256, 0, stream>>>(b->ncols, b->B, nt);
    // outputs never reached stay NaN (all-ones bit pattern)
    DSB_CUDA(cudaMemsetAsync(ys_dev, 0xFF, (size_t)nt * b->prob.nout * b->B * 8, stream));
    b->last_launches = 0;
    DSB_CUDA(cudaEventRecord(b->ev0, stream));
    const PluginModel* pm = plugin_of(b->prob.model);
    if (!pm && (b->prob.model < 0 || b->prob.model >= DSB_MODEL_COUNT)) return fail(DSB_BAD_ARG, "unknown model id");
    std::vector<double> atol_full((size_t)b->prob.n);
    for (int i = 0; i < b->prob.n; ++i) atol_full[i] = b->prob.atol.size() == 1 ? b->prob.atol[0] : b->prob.atol[i];
    if (const char* q = getenv("DSB_EXEC_MODE")) b->coop.exec_mode = atoi(q);
    b->coop.color_host = color_full.empty() ? nullptr : color_full.data();
    b->coop.nz_host = nz_full.empty() ? nullptr : nz_full.data();      // test hook: 1 = lane kernels, 2 = cooperative
    b->coop.ys_im = ys_im; b->coop.ys_im_used = nullptr;
    if (wrote_im) *wrote_im = 0;
    const dsb_launch_fn launch = pm ? pm->launch : g_launch_table[b->prob.model];
    cudaError_t lerr = launch(&pa, &bb, method, stream, b->ev_mid, b->work_counter, &b->coop, atol_full.data(), &b->last_launches);
    if (lerr == cudaErrorNotSupported && pa.ragged)
        return fail(DSB_ERR, "solve(final_time) is built for the thread-per-instance kernels: n <= 16, no reset function, no sensitivities");
    if (lerr == cudaErrorNotSupported && b->prob.sens)
        return fail(DSB_ERR, "forward sensitivities are built for equation sets with sens_mul / init_sens, no root / output / reset function and n <= 16");
    if (lerr == cudaErrorNotSupported) return fail(DSB_ERR, "this execution mode is not available for this equation set and method (thread per instance: n <= 16; banded thread per instance: component-wise equations with a declared band, n > 16; banded warp per instance: the same, BDF, no reset function; block per instance: n <= 512)");
    if (lerr != cudaSuccess) return fail(DSB_ERR, std::string("kernel launch: ") + cudaGetErrorString(lerr));
    if (b->coop.ys_im_used) {
        // the kernel wrote the instance-major layout (warp-per-instance banded kernel)
        if (ys_im && b->coop.ys_im_used == ys_im && wrote_im) {
            *wrote_im = 1;
        } else {
            const int m = nt * b->prob.nout;
            dim3 grid((unsigned)((b->B + 31) / 32), (unsigned)((m + 31) / 32)), block(32, 8);
            dsb_im_to_batch_major_kernel<<<grid, block, 0, stream>>>(b->coop.ys_im_used, ys_dev, b->B, m);
            DSB_CUDA(cudaGetLastError());
            b->last_launches += 1;
        }
    }
    DSB_CUDA(cudaEventRecord(b->ev1, stream));
    b->have_timing = true;
    return DSB_OK;
}

int dsb_batch_solve_dense(dsb_batch* b, int32_t method, const double* t_eval, int32_t nt, double* ys_dev, void* stream) {
    return solve_impl(b, method, t_eval, nt, ys_dev, stream, 0);
}
int dsb_batch_step_and_interpolate(dsb_batch* b, int32_t method, const double* t_points, int32_t npts, double* ys_dev,
                                   void* stream) {
    return solve_impl(b, method, t_points, npts, ys_dev, stream, 1);
}

int dsb_batch_solve_dense_sensitivities(dsb_batch* b, int32_t method, const double* t_eval, int32_t nt, double* ys_dev, double* sens_dev,
                                        void* stream) {
    if (!sens_dev) return fail(DSB_BAD_ARG, "sens_dev is NULL");
    SolveExtras ex; ex.sens_dev = sens_dev;
    return solve_impl(b, method, t_eval, nt, ys_dev, stream, 0, nullptr, nullptr, &ex);
}
int dsb_batch_step_and_interpolate_sensitivities(dsb_batch* b, int32_t method, const double* t_points, int32_t npts, double* ys_dev,
                                                 double* sens_dev, void* stream) {
    if (!sens_dev) return fail(DSB_BAD_ARG, "sens_dev is NULL");
    SolveExtras ex; ex.sens_dev = sens_dev;
    return solve_impl(b, method, t_points, npts, ys_dev, stream, 1, nullptr, nullptr, &ex);
}

// ---- OdeSolverMethod::solve(final_time): ragged results in two passes -------------------------------------------------------
static int ragged_pass(dsb_batch* b, int32_t method, double final_time, const SolveExtras& ex, cudaStream_t stream) {
    const size_t dummy = (size_t)b->prob.nout * b->B * 8;           // the dense result block of a one-point solve: unused
    if (b->ys_own_bytes < dummy) {
        cudaFree(b->ys_own); b->ys_own = nullptr; b->ys_own_bytes = 0;
        DSB_CUDA(cudaMalloc((void**)&b->ys_own, dummy));
        b->ys_own_bytes = dummy;
    }
    return solve_impl(b, method, &final_time, 1, b->ys_own, stream, 0, nullptr, nullptr, &ex);
}
int dsb_batch_solve_count(dsb_batch* b, int32_t method, double final_time, int64_t* total_columns) {
    if (!b || !total_columns) return fail(DSB_BAD_ARG, "NULL argument");
    if (b->prob.sens) return fail(DSB_BAD_ARG, "solve(final_time) is not built for problems with sensitivities");
    DSB_CUDA(cudaSetDevice(b->device));
    SolveExtras ex; ex.ragged = 1;
    int rc = ragged_pass(b, method, final_time, ex, 0);
    if (rc != DSB_OK) return rc;
    std::vector<int32_t> nc((size_t)b->B);
    DSB_CUDA(cudaMemcpy(nc.data(), b->ncols, (size_t)b->B * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> off((size_t)b->B + 1);
    off[0] = 0;
    for (int64_t k = 0; k < b->B; ++k) off[(size_t)k + 1] = off[(size_t)k] + nc[(size_t)k];
    if (!b->rag_off) DSB_CUDA(cudaMalloc((void**)&b->rag_off, ((size_t)b->B + 1) * 8));
    DSB_CUDA(cudaMemcpy(b->rag_off, off.data(), ((size_t)b->B + 1) * 8, cudaMemcpyHostToDevice));
    b->rag_total = off[(size_t)b->B]; b->rag_method = method; b->rag_final_time = final_time;
    *total_columns = b->rag_total;
    return DSB_OK;
}
int dsb_batch_solve_offsets(dsb_batch* b, int64_t* offsets_host) {
    if (!b || !offsets_host) return fail(DSB_BAD_ARG, "NULL argument");
    if (b->rag_total < 0) return fail(DSB_BAD_ARG, "dsb_batch_solve_count has not run on this batch");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaMemcpy(offsets_host, b->rag_off, ((size_t)b->B + 1) * 8, cudaMemcpyDeviceToHost));
    return DSB_OK;
}
int dsb_batch_solve_write(dsb_batch* b, int32_t method, double final_time, double* ts_dev, double* ys_dev, void* stream) {
    if (!b || !ts_dev || !ys_dev) return fail(DSB_BAD_ARG, "NULL argument");
    if (b->rag_total < 0 || b->rag_method != method || b->rag_final_time != final_time)
        return fail(DSB_BAD_ARG, "dsb_batch_solve_write follows a dsb_batch_solve_count with the same method and final time");
    DSB_CUDA(cudaSetDevice(b->device));
    SolveExtras ex; ex.ragged = 2; ex.rag_off = b->rag_off; ex.rag_ts = ts_dev; ex.rag_ys = ys_dev;
    return ragged_pass(b, method, final_time, ex, (cudaStream_t)stream);
}
int dsb_batch_solve_write_host(dsb_batch* b, int32_t method, double final_time, double* ts_host, double* ys_host) {
    if (!b || !ts_host || !ys_host) return fail(DSB_BAD_ARG, "NULL argument");
    if (b->rag_total < 0) return fail(DSB_BAD_ARG, "dsb_batch_solve_count has not run on this batch");
    DSB_CUDA(cudaSetDevice(b->device));
    const size_t tb = (size_t)(b->rag_total > 0 ? b->rag_total : 1) * 8, yb = tb * (size_t)b->prob.nout;
    if (ensure_stage(b, tb + yb) != DSB_OK) return DSB_ERR;
    double* ts_dev = (double*)b->stage;
    double* ys_dev = (double*)((char*)b->stage + tb);
    int rc = dsb_batch_solve_write(b, method, final_time, ts_dev, ys_dev, nullptr);
    if (rc != DSB_OK) return rc;
    DSB_CUDA(cudaMemcpy(ts_host, ts_dev, (size_t)b->rag_total * 8, cudaMemcpyDeviceToHost));
    DSB_CUDA(cudaMemcpy(ys_host, ys_dev, (size_t)b->rag_total * 8 * (size_t)b->prob.nout, cudaMemcpyDeviceToHost));
    return DSB_OK;
}

int dsb_batch_set_execution(dsb_batch* b, int32_t mode) {
    if (!b || mode < 0 || mode > 4) return fail(DSB_BAD_ARG, "mode must be 0 (automatic), 1 (thread per instance), 2 (block per instance), 3 (thread per instance, banded, state in global memory) or 4 (warp per instance, banded, state in shared memory)");
    b->coop.exec_mode = mode;
    return DSB_OK;
}

int dsb_batch_get_stats_device(dsb_batch* b, int64_t* stats_dev, void* stream) {
    if (!b || !stats_dev) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    const int64_t total = b->B * DSB_NSTATS;
    dsb_stats_to_host_layout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        b->stats, stats_dev, b->B, b->sparsity_probe_jac_muls);
    DSB_CUDA(cudaGetLastError());
    return DSB_OK;
}

int dsb_batch_get_stats(dsb_batch* b, int64_t* stats_host) {
    if (!b || !stats_host) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    const size_t bytes = (size_t)b->B * DSB_NSTATS * 8;
    if (ensure_stage(b, bytes) != DSB_OK) return DSB_ERR;
    DSB_CUDA(cudaDeviceSynchronize());
    if (dsb_batch_get_stats_device(b, (int64_t*)b->stage, nullptr) != DSB_OK) return DSB_ERR;
    DSB_CUDA(cudaMemcpy(stats_host, b->stage, bytes, cudaMemcpyDeviceToHost));
    return DSB_OK;
}
int dsb_batch_get_status(dsb_batch* b, int32_t* status_host) {
    if (!b || !status_host) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaDeviceSynchronize());
    DSB_CUDA(cudaMemcpy(status_host, b->status, (size_t)b->B * 4, cudaMemcpyDeviceToHost));
    return DSB_OK;
}
int dsb_batch_get_final_state(dsb_batch* b, double* t_host, double* h_host, int32_t* order_host) {
    if (!b) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaDeviceSynchronize());
    if (t_host) DSB_CUDA(cudaMemcpy(t_host, b->fin_t, (size_t)b->B * 8, cudaMemcpyDeviceToHost));
    if (h_host) DSB_CUDA(cudaMemcpy(h_host, b->fin_h, (size_t)b->B * 8, cudaMemcpyDeviceToHost));
    if (order_host) DSB_CUDA(cudaMemcpy(order_host, b->fin_order, (size_t)b->B * 4, cudaMemcpyDeviceToHost));
    return DSB_OK;
}
int dsb_batch_get_root_info(dsb_batch* b, int32_t* root_idx_host, int32_t* ncols_host) {
    if (!b) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaDeviceSynchronize());
    if (root_idx_host) DSB_CUDA(cudaMemcpy(root_idx_host, b->root_idx, (size_t)b->B * 4, cudaMemcpyDeviceToHost));
    if (ncols_host) DSB_CUDA(cudaMemcpy(ncols_host, b->ncols, (size_t)b->B * 4, cudaMemcpyDeviceToHost));
    return DSB_OK;
}
int dsb_batch_device_views(dsb_batch* b, const int32_t** stats_dev, const int32_t** status_dev) {
    if (!b) return fail(DSB_BAD_ARG, "NULL argument");
    if (stats_dev) *stats_dev = b->stats;
    if (status_dev) *status_dev = b->status;
    return DSB_OK;
}

int dsb_batch_sum_stat(dsb_batch* b, int32_t stat, int64_t* total) {
    if (!b || !total || stat < 0 || stat >= DSB_NSTATS) return fail(DSB_BAD_ARG, "bad argument");
    DSB_CUDA(cudaSetDevice(b->device));
    if (ensure_stage(b, 8) != DSB_OK) return DSB_ERR;
    DSB_CUDA(cudaDeviceSynchronize());
    DSB_CUDA(cudaMemset(b->stage, 0, 8));
    dsb_sum_stat_kernel<<<296, 256>>>(b->stats + (size_t)stat * b->B, b->B, (unsigned long long*)b->stage);
    DSB_CUDA(cudaGetLastError());
    unsigned long long v = 0;
    DSB_CUDA(cudaMemcpy(&v, b->stage, 8, cudaMemcpyDeviceToHost));
    if (stat == DSB_STAT_RHS_JAC_MULS) v += (unsigned long long)b->sparsity_probe_jac_muls * (unsigned long long)b->B;
    *total = (int64_t)v;
    return DSB_OK;
}

int dsb_batch_last_kernel_ms(dsb_batch* b, float* ms) {
    if (!b || !ms || !b->have_timing) return fail(DSB_BAD_ARG, "no solve has been timed");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaEventSynchronize(b->ev1));
    DSB_CUDA(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return DSB_OK;
}
int dsb_batch_last_integrator_ms(dsb_batch* b, float* ms) {
    if (!b || !ms || !b->have_timing) return fail(DSB_BAD_ARG, "no solve has been timed");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaEventSynchronize(b->ev1));
    DSB_CUDA(cudaEventElapsedTime(ms, b->ev_mid, b->ev1));
    return DSB_OK;
}
int dsb_batch_last_launch_count(dsb_batch* b, int32_t* launches) {
    if (!b || !launches) return fail(DSB_BAD_ARG, "NULL argument");
    *launches = b->last_launches;
    return DSB_OK;
}

int dsb_batch_debug_words(dsb_batch* b, uint64_t* words_host) {
    if (!b || !words_host) return fail(DSB_BAD_ARG, "NULL argument");
    DSB_CUDA(cudaSetDevice(b->device));
    DSB_CUDA(cudaDeviceSynchronize());
    DSB_CUDA(cudaMemcpy(words_host, b->work_counter + 1, 31 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DSB_OK;
}

// One chunk (or the whole batch) of a *_host call on `stream`: parameters up, kernels, results down.  Nothing here waits
// for the device; the caller synchronises.  The instance-major staging block doubles as the kernels' result block when
// they write that layout (warp-per-instance banded kernel).
static int solve_host_enqueue(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams, const double* t_eval,
                              int32_t nt, double* ys_host, int64_t* stats_host, int32_t* status_host, int free_running,
                              cudaStream_t stream) {
    const int n = b->prob.nout;                       // rows of a result column
    const size_t ys_bytes = (size_t)nt * n * b->B * 8;
    size_t stage_need = ys_bytes + (size_t)b->B * DSB_NSTATS * 8;
    if ((size_t)b->B * (nparams > 0 ? nparams : 1) * 8 > stage_need) stage_need = (size_t)b->B * nparams * 8;
    if (ensure_stage(b, stage_need) != DSB_OK) return DSB_ERR;
    if (b->ys_own_bytes < ys_bytes) {
        cudaFree(b->ys_own); b->ys_own = nullptr; b->ys_own_bytes = 0;
        DSB_CUDA(cudaMalloc((void**)&b->ys_own, ys_bytes));
        b->ys_own_bytes = ys_bytes;
    }
    int extra = 0;
    if (nparams > 0) {
        DSB_CUDA(cudaMemcpyAsync(b->stage, params_host, (size_t)b->B * nparams * 8, cudaMemcpyHostToDevice, stream));
        int rc = dsb_batch_set_params_device(b, (const double*)b->stage, b->B, nparams, stream);
        if (rc != DSB_OK) return rc;
        ++extra;
    }
    int wrote_im = 0;
    int rc = solve_impl(b, method, t_eval, nt, b->ys_own, stream, free_running, (double*)b->stage, &wrote_im);
    if (rc != DSB_OK) return rc;
    if (!wrote_im) {
        const int m = nt * n;
        dim3 grid((unsigned)((b->B + 31) / 32), (unsigned)((m + 31) / 32)), block(32, 8);
        dsb_to_instance_major_kernel<<<grid, block, 0, stream>>>(b->ys_own, (double*)b->stage, b->B, m);
        DSB_CUDA(cudaGetLastError());
        ++extra;
    }
    DSB_CUDA(cudaMemcpyAsync(ys_host, b->stage, ys_bytes, cudaMemcpyDeviceToHost, stream));
    if (stats_host) {
        int64_t* stats_stage = (int64_t*)((char*)b->stage + ys_bytes);
        if (dsb_batch_get_stats_device(b, stats_stage, stream) != DSB_OK) return DSB_ERR;
        ++extra;
        DSB_CUDA(cudaMemcpyAsync(stats_host, stats_stage, (size_t)b->B * DSB_NSTATS * 8, cudaMemcpyDeviceToHost, stream));
    }
    if (status_host) DSB_CUDA(cudaMemcpyAsync(status_host, b->status, (size_t)b->B * 4, cudaMemcpyDeviceToHost, stream));
    b->last_launches += extra;
    return DSB_OK;
}

static int solve_host_impl(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                           const double* t_eval, int32_t nt, double* ys_host, int64_t* stats_host,
                           int32_t* status_host, int free_running) {
    if (!b || !ys_host) return fail(DSB_BAD_ARG, "NULL argument");
    if (!t_eval || nt < 1) return fail(DSB_BAD_ARG, "t_eval must hold at least one time");
    if (method != DSB_METHOD_BDF && method != DSB_METHOD_TR_BDF2 && method != DSB_METHOD_ESDIRK34)
        return fail(DSB_BAD_ARG, "unknown method");
    if (nparams > 0 && (!params_host || nparams != b->prob.np)) return fail(DSB_BAD_ARG, "parameter shape mismatch");
    DSB_CUDA(cudaSetDevice(b->device));
    // OPTIONAL pipelining (DSB_HOST_CHUNKS = 2..16, default off): the batch runs as child batches on their own streams --
    // the persistent kernels of consecutive chunks hand the SMs over as their work counters run out, and the copies of
    // the other chunks run underneath (host buffers should be pinned for that).  Measured on one B200, 10^6 Robertson
    // instances (24 MB up, 148 MB down, pinned): 62.4 ms direct, 65.1 ms in 4 chunks -- each chunk's kernel pays its own
    // tail of late-finishing instances, which costs more than the 3 ms of copies it hides -- hence off by default; it is
    // there for hosts whose device -> host path is slow or shared (8 ranks on one PCIe root).
    int nchunks = 1;
    if (const char* q = getenv("DSB_HOST_CHUNKS")) nchunks = atoi(q);
    if (nchunks > 16) nchunks = 16;
    if (nchunks < 2 || b->B < (int64_t)nchunks * 65536) {
        int rc = solve_host_enqueue(b, method, params_host, nparams, t_eval, nt, ys_host, stats_host, status_host, free_running, 0);
        if (rc != DSB_OK) return rc;
        DSB_CUDA(cudaStreamSynchronize(0));
        return DSB_OK;
    }
    const int n = b->prob.nout;
    if ((int)b->chunks.size() != nchunks) {
        for (dsb_batch* c : b->chunks) dsb_batch_free(c);
        for (cudaStream_t st : b->chunk_streams) cudaStreamDestroy(st);
        for (cudaEvent_t ev : b->chunk_done) cudaEventDestroy(ev);
        b->chunks.clear(); b->chunk_streams.clear(); b->chunk_done.clear();
        for (int k = 0; k < nchunks; ++k) {
            const int64_t k0 = b->B * k / nchunks, k1 = b->B * (k + 1) / nchunks;
            dsb_batch* c = nullptr;
            if (dsb_batch_new(&b->prob, k1 - k0, b->device, &c) != DSB_OK) return DSB_ERR;
            b->chunks.push_back(c);
            cudaStream_t st; cudaEvent_t ev;
            DSB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            DSB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            b->chunk_streams.push_back(st); b->chunk_done.push_back(ev);
        }
    }
    DSB_CUDA(cudaEventRecord(b->ev0, 0));
    DSB_CUDA(cudaEventRecord(b->ev_mid, 0));
    b->last_launches = 0;
    for (int k = 0; k < nchunks; ++k) {
        dsb_batch* c = b->chunks[(size_t)k];
        cudaStream_t st = b->chunk_streams[(size_t)k];
        const int64_t k0 = b->B * k / nchunks;
        c->prob = b->prob; c->coop.exec_mode = b->coop.exec_mode;
        DSB_CUDA(cudaStreamWaitEvent(st, b->ev0, 0));
        if (nparams == 0 && b->prob.np > 0)        // parameters set earlier on the parent: hand the chunk its columns
            for (int j = 0; j < b->prob.np; ++j)
                DSB_CUDA(cudaMemcpyAsync(c->params + (size_t)j * c->B, b->params + (size_t)j * b->B + k0, (size_t)c->B * 8, cudaMemcpyDeviceToDevice, st));
        int rc = solve_host_enqueue(c, method, params_host ? params_host + (size_t)k0 * nparams : nullptr, nparams, t_eval, nt,
                                    ys_host + (size_t)k0 * nt * n, stats_host ? stats_host + (size_t)k0 * DSB_NSTATS : nullptr,
                                    status_host ? status_host + k0 : nullptr, free_running, st);
        if (rc != DSB_OK) return rc;
        // the parent's per-instance arrays (read by the getters after the call) take the chunk's rows
        DSB_CUDA(cudaMemcpy2DAsync(b->stats + k0, (size_t)b->B * 4, c->stats, (size_t)c->B * 4, (size_t)c->B * 4, DSB_NSTATS, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->status + k0, c->status, (size_t)c->B * 4, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->fin_t + k0, c->fin_t, (size_t)c->B * 8, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->fin_h + k0, c->fin_h, (size_t)c->B * 8, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->fin_order + k0, c->fin_order, (size_t)c->B * 4, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->root_idx + k0, c->root_idx, (size_t)c->B * 4, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaMemcpyAsync(b->ncols + k0, c->ncols, (size_t)c->B * 4, cudaMemcpyDeviceToDevice, st));
        DSB_CUDA(cudaEventRecord(b->chunk_done[(size_t)k], st));
        DSB_CUDA(cudaStreamWaitEvent(0, b->chunk_done[(size_t)k], 0));
        b->last_launches += c->last_launches;
        b->sparsity_probe_jac_muls = c->sparsity_probe_jac_muls;
    }
    DSB_CUDA(cudaEventRecord(b->ev1, 0));
    b->have_timing = true;
    DSB_CUDA(cudaStreamSynchronize(0));
    return DSB_OK;
}

// Sensitivities with HOST buffers: parameters up, solve, the states and the sensitivities transposed to instance-major
// through the staging block (one after the other on the default stream), counters and status down, synchronise.
static int solve_sens_host_impl(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams, const double* t_eval,
                                int32_t nt, double* ys_host, double* sens_host, int64_t* stats_host, int32_t* status_host, int free_running) {
    if (!b || !ys_host || !sens_host) return fail(DSB_BAD_ARG, "NULL argument");
    if (!t_eval || nt < 1) return fail(DSB_BAD_ARG, "t_eval must hold at least one time");
    if (nparams < 1 || !params_host || nparams != b->prob.np) return fail(DSB_BAD_ARG, "parameter shape mismatch");
    if (!b->prob.sens) return fail(DSB_BAD_ARG, "the problem has no sensitivities enabled (dsb_problem_set_sensitivities)");
    DSB_CUDA(cudaSetDevice(b->device));
    cudaStream_t stream = 0;
    const int n = b->prob.n;
    const size_t ys_bytes = (size_t)nt * n * b->B * 8, sens_bytes = ys_bytes * (size_t)nparams;
    size_t stage_need = sens_bytes;
    if ((size_t)b->B * DSB_NSTATS * 8 > stage_need) stage_need = (size_t)b->B * DSB_NSTATS * 8;
    if ((size_t)b->B * nparams * 8 > stage_need) stage_need = (size_t)b->B * nparams * 8;
    if (ensure_stage(b, stage_need) != DSB_OK) return DSB_ERR;
    if (b->ys_own_bytes < ys_bytes) {
        cudaFree(b->ys_own); b->ys_own = nullptr; b->ys_own_bytes = 0;
        DSB_CUDA(cudaMalloc((void**)&b->ys_own, ys_bytes));
        b->ys_own_bytes = ys_bytes;
    }
    if (b->sens_own_bytes < sens_bytes) {
        cudaFree(b->sens_own); b->sens_own = nullptr; b->sens_own_bytes = 0;
        DSB_CUDA(cudaMalloc((void**)&b->sens_own, sens_bytes));
        b->sens_own_bytes = sens_bytes;
    }
    DSB_CUDA(cudaMemcpyAsync(b->stage, params_host, (size_t)b->B * nparams * 8, cudaMemcpyHostToDevice, stream));
    int rc = dsb_batch_set_params_device(b, (const double*)b->stage, b->B, nparams, stream);
    if (rc != DSB_OK) return rc;
    SolveExtras ex; ex.sens_dev = b->sens_own;
    rc = solve_impl(b, method, t_eval, nt, b->ys_own, stream, free_running, nullptr, nullptr, &ex);
    if (rc != DSB_OK) return rc;
    dim3 block(32, 8);
    {
        const int m = nt * n;
        dim3 grid((unsigned)((b->B + 31) / 32), (unsigned)((m + 31) / 32));
        dsb_to_instance_major_kernel<<<grid, block, 0, stream>>>(b->ys_own, (double*)b->stage, b->B, m);
        DSB_CUDA(cudaGetLastError());
        DSB_CUDA(cudaMemcpyAsync(ys_host, b->stage, ys_bytes, cudaMemcpyDeviceToHost, stream));
    }
    {
        const int m = nt * nparams * n;
        dim3 grid((unsigned)((b->B + 31) / 32), (unsigned)((m + 31) / 32));
        dsb_to_instance_major_kernel<<<grid, block, 0, stream>>>(b->sens_own, (double*)b->stage, b->B, m);
        DSB_CUDA(cudaGetLastError());
        DSB_CUDA(cudaMemcpyAsync(sens_host, b->stage, sens_bytes, cudaMemcpyDeviceToHost, stream));
    }
    int extra = 3;
    if (stats_host) {
        if (dsb_batch_get_stats_device(b, (int64_t*)b->stage, stream) != DSB_OK) return DSB_ERR;
        ++extra;
        DSB_CUDA(cudaMemcpyAsync(stats_host, b->stage, (size_t)b->B * DSB_NSTATS * 8, cudaMemcpyDeviceToHost, stream));
    }
    if (status_host) DSB_CUDA(cudaMemcpyAsync(status_host, b->status, (size_t)b->B * 4, cudaMemcpyDeviceToHost, stream));
    b->last_launches += extra;
    DSB_CUDA(cudaStreamSynchronize(stream));
    return DSB_OK;
}

int dsb_batch_solve_dense_sensitivities_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams, const double* t_eval,
                                             int32_t nt, double* ys_host, double* sens_host, int64_t* stats_host, int32_t* status_host) {
    return solve_sens_host_impl(b, method, params_host, nparams, t_eval, nt, ys_host, sens_host, stats_host, status_host, 0);
}
int dsb_batch_step_and_interpolate_sensitivities_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                                                      const double* t_points, int32_t npts, double* ys_host, double* sens_host,
                                                      int64_t* stats_host, int32_t* status_host) {
    return solve_sens_host_impl(b, method, params_host, nparams, t_points, npts, ys_host, sens_host, stats_host, status_host, 1);
}

int dsb_batch_solve_dense_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                               const double* t_eval, int32_t nt, double* ys_host, int64_t* stats_host,
                               int32_t* status_host) {
    return solve_host_impl(b, method, params_host, nparams, t_eval, nt, ys_host, stats_host, status_host, 0);
}
int dsb_batch_step_and_interpolate_host(dsb_batch* b, int32_t method, const double* params_host, int32_t nparams,
                                        const double* t_points, int32_t npts, double* ys_host, int64_t* stats_host,
                                        int32_t* status_host) {
    return solve_host_impl(b, method, params_host, nparams, t_points, npts, ys_host, stats_host, status_host, 1);
}

int dsb_lu_factor_batched(double* a_dev, int32_t n, int64_t nbatch, int32_t* piv_dev, int32_t* info_dev, void* stream) {
    if (!a_dev || !piv_dev || !info_dev || n < 1 || nbatch < 1) return fail(DSB_BAD_ARG, "bad argument to dsb_lu_factor_batched");
    cudaError_t e = dsb_launch_lu_factor(a_dev, n, nbatch, piv_dev, info_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DSB_ERR, std::string("dsb_lu_factor_batched: ") + cudaGetErrorString(e));
    return DSB_OK;
}
int dsb_lu_solve_batched(const double* lu_dev, const int32_t* piv_dev, double* b_dev, int32_t n, int64_t nbatch,
                         int32_t* info_dev, void* stream) {
    if (!lu_dev || !piv_dev || !b_dev || !info_dev || n < 1 || nbatch < 1) return fail(DSB_BAD_ARG, "bad argument to dsb_lu_solve_batched");
    cudaError_t e = dsb_launch_lu_solve(lu_dev, piv_dev, b_dev, n, nbatch, info_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DSB_ERR, std::string("dsb_lu_solve_batched: ") + cudaGetErrorString(e));
    return DSB_OK;
}

int dsb_lu_factor_instance_major(double* a_dev, int32_t n, int64_t nbatch, int32_t* piv_dev, int32_t* info_dev, void* stream) {
    if (!a_dev || !piv_dev || !info_dev || n < 1 || nbatch < 1) return fail(DSB_BAD_ARG, "bad argument to dsb_lu_factor_instance_major");
    cudaError_t e = dsb_launch_lu_factor_im(a_dev, n, nbatch, piv_dev, info_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DSB_ERR, std::string("dsb_lu_factor_instance_major: ") + cudaGetErrorString(e));
    return DSB_OK;
}
int dsb_lu_solve_instance_major(const double* lu_dev, const int32_t* piv_dev, double* b_dev, int32_t n, int64_t nbatch,
                                int32_t* info_dev, void* stream) {
    if (!lu_dev || !piv_dev || !b_dev || !info_dev || n < 1 || nbatch < 1) return fail(DSB_BAD_ARG, "bad argument to dsb_lu_solve_instance_major");
    cudaError_t e = dsb_launch_lu_solve_im(lu_dev, piv_dev, b_dev, n, nbatch, info_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DSB_ERR, std::string("dsb_lu_solve_instance_major: ") + cudaGetErrorString(e));
    return DSB_OK;
}

}  // extern "C"
