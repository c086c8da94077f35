// rbq_api.cu — the C ABI of include/rbq.h: handle lifetime, argument checking, host<->device staging.
// No torch types, no exceptions across the boundary, no CPU fallback: every compute entry point
// ends in a kernel launch on the handle's device or returns an error code.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <new>

#include "rbq_internal.h"
#include "rbq_models.cuh"

using namespace rbq;

// dbuf roles
enum { B_PULSE = 0, B_PARTIALS, B_SCALARS, B_IN0, B_IN1, B_IN2, B_IN3, B_IN4, B_OUT0, B_OUT1, B_STATS, B_IN5, B_IDX, B_QUAT, B_NOISE };

static int fail(rbq_handle* h, int code, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(h, (int)e_, "%s: %s", #call, cudaGetErrorString(e_));       \
    } while (0)
#define CKL(expr)                                                                                      \
    do {                                                                                               \
        int r_ = (expr);                                                                               \
        if (r_ < 0) return fail(h, r_, "%s: argument error %d", #expr, r_);                            \
        if (r_ > 0) return fail(h, r_, "%s: %s", #expr, cudaGetErrorString((cudaError_t)r_));          \
    } while (0)

static int dev_buf(rbq_handle* h, int role, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (h->dcap[role] < bytes) {
        if (h->dbuf[role]) {
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaFree(h->dbuf[role]));
            h->dbuf[role] = nullptr; h->dcap[role] = 0;
        }
        size_t cap = bytes + bytes / 4;
        cap = (cap + 255) & ~(size_t)255;
        CK(cudaMalloc(&h->dbuf[role], cap));
        h->dcap[role] = cap;
    }
    *out = h->dbuf[role];
    return 0;
}
int rbq_scratch(rbq_handle* h, int role, size_t bytes, void** out) { return dev_buf(h, role, bytes, out); }
#define ENTER(h)                                                \
    if (!(h)) return RBQ_E_NULL;                                \
    (h)->err[0] = 0;                                            \
    CK(cudaSetDevice((h)->device))

// ---- gates (spin.jl:301-323) ---------------------------------------------------------------------------
typedef std::complex<double> cplx;
static bool gate_matrix(int gate, cplx g[2][2]) {
    const double r = 1.0 / sqrt(2.0);
    const cplx I(0, 1);
    switch (gate) {
    case RBQ_GATE_ZPIBY2: g[0][0] = (1.0 - I) * r; g[0][1] = 0; g[1][0] = 0; g[1][1] = (1.0 + I) * r; return true;
    case RBQ_GATE_YPIBY2: g[0][0] = r; g[0][1] = -r; g[1][0] = r; g[1][1] = r; return true;
    case RBQ_GATE_XPIBY2: g[0][0] = r; g[0][1] = -I * r; g[1][0] = -I * r; g[1][1] = r; return true;
    case RBQ_GATE_XPI: g[0][0] = 0; g[0][1] = -I; g[1][0] = -I; g[1][1] = 0; return true;
    }
    return false;
}
static void mul2(const cplx a[2][2], const cplx b[2][2], cplx c[2][2]) {
    cplx t[2][2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) c[i][j] = t[i][j];
}
// real isomorphism of rbqoc.jl:204-210, built from the 2x2 power directly (no 4x4 products)
static void iso_of(const cplx g[2][2], double* G) {
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            G[4 * i + j] = g[i][j].real(); G[4 * i + j + 2] = -g[i][j].imag();
            G[4 * (i + 2) + j] = g[i][j].imag(); G[4 * (i + 2) + j + 2] = g[i][j].real();
        }
}
// R_ij = tr(sigma_i g sigma_j g^dagger)/2 : how g rho g^dagger moves the Bloch vector
static void rot_of(const cplx g[2][2], double* R) {
    const cplx I(0, 1);
    const cplx sg[3][2][2] = {{{0, 1}, {1, 0}}, {{0, -I}, {I, 0}}, {{1, 0}, {0, -1}}};
    cplx gd[2][2] = {{std::conj(g[0][0]), std::conj(g[1][0])}, {std::conj(g[0][1]), std::conj(g[1][1])}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            cplx a[2][2], b[2][2];
            mul2(g, sg[j], a); mul2(a, gd, a); mul2(sg[i], a, b);
            R[3 * i + j] = 0.5 * (b[0][0] + b[1][1]).real();
        }
}
static bool make_gate_tables(int gate, GateIso* gi, GateRot* gr) {
    cplx g[3][2][2];
    if (!gate_matrix(gate, g[0])) return false;
    mul2(g[0], g[0], g[1]);
    mul2(g[1], g[0], g[2]);
    for (int k = 0; k < 3; ++k) {
        if (gi) iso_of(g[k], gi->g[k]);
        if (gr) rot_of(g[k], gr->r[k]);
    }
    return true;
}

// ---- tuning options: latched from the environment in rbq_create, changed later only by rbq_set_option ----------------
static const char* const kOptionNames[] = {"RBQ_NO_TABLE_CLASS", "RBQ_MC_ZEROCOPY", "RBQ_MC_CHUNK_WAVES", "RBQ_MC_VARIANT", "RBQ_PEAK_MODE",
                                           "RBQ_NOISE_GATE_PARALLEL", "RBQ_NOISE_MATERIALISE", "RBQ_LINDBLAD_NO_SHARED_MAP", "RBQ_LINDBLAD_TABLE", "RBQ_LINDBLAD_SPT",
                                           "RBQ_ROLLOUT_PACKED", "RBQ_ROLLOUT_SCAN", "RBQ_JAC_NT_UNSCENTED", "RBQ_JAC_STRUCTURED",
                                           "RBQ_JAC_STAGE_KB", "RBQ_JAC_STAGED", "RBQ_BP_TMA", "RBQ_BP_GENERIC", "RBQ_BP_WARPS", "RBQ_BP_CLUSTER"};
int rbq_apply_option(rbq_handle* h, const char* name, const char* value) {
    const bool unset = !value || !*value;
    const int v = unset ? -1 : atoi(value);
    auto is = [&](const char* n) { return strcmp(name, n) == 0; };
    if (is("RBQ_NO_TABLE_CLASS")) h->opt.no_table_class = unset ? 0 : (v != 0);
    else if (is("RBQ_MC_ZEROCOPY")) h->opt.mc_zerocopy = unset ? 1 : (v != 0);
    else if (is("RBQ_MC_CHUNK_WAVES")) {
        int a1 = 0, a2 = 0;
        h->opt.chunk_first = 1; h->opt.chunk_rest = 1;   // in waves (measured best: tools/tune_e2e.py)
        if (!unset) {
            if (sscanf(value, "%d,%d", &a1, &a2) != 2 || a1 < 1 || a2 < 1) return RBQ_E_SIZE;
            h->opt.chunk_first = a1; h->opt.chunk_rest = a2;
        }
    }
    else if (is("RBQ_MC_VARIANT")) h->opt.mc_variant = v;
    else if (is("RBQ_NOISE_MATERIALISE")) h->opt.noise_materialise = unset ? 1 : (v != 0);
    else if (is("RBQ_PEAK_MODE")) h->opt.peak_mode = unset ? 0 : v;
    else if (is("RBQ_NOISE_GATE_PARALLEL")) h->opt.noise_gate_parallel = unset ? -1 : (v != 0);
    else if (is("RBQ_LINDBLAD_NO_SHARED_MAP")) h->opt.lindblad_shared_map = unset ? 1 : (v == 0);
    else if (is("RBQ_LINDBLAD_TABLE")) h->opt.lindblad_table = unset ? 1 : (v != 0);
    else if (is("RBQ_LINDBLAD_SPT")) h->opt.lindblad_spt = (v == 1 || v == 2) ? v : -1;
    else if (is("RBQ_ROLLOUT_PACKED")) h->opt.rollout_packed = unset ? -1 : (v != 0);
    else if (is("RBQ_ROLLOUT_SCAN")) h->opt.rollout_scan = unset ? -1 : (v != 0);
    else if (is("RBQ_JAC_NT_UNSCENTED")) h->opt.jac_nt_unscented = (v == 2) ? 2 : 1;
    else if (is("RBQ_JAC_STRUCTURED")) h->opt.jac_structured = unset ? 1 : (v != 0);
    else if (is("RBQ_JAC_STAGE_KB")) h->opt.jac_stage_kb = (v >= 1 && v <= 200) ? v : -1;
    else if (is("RBQ_JAC_STAGED")) h->opt.jac_staged = unset ? -1 : (v != 0);
    else if (is("RBQ_BP_TMA")) h->opt.bp_tma = v;
    else if (is("RBQ_BP_GENERIC")) h->opt.bp_generic = unset ? 0 : (v != 0);
    else if (is("RBQ_BP_WARPS")) h->opt.bp_warps = v;
    else if (is("RBQ_BP_CLUSTER")) h->opt.bp_cluster = unset ? 0 : v;
    else return RBQ_E_NULL;
    return RBQ_OK;
}

extern "C" {

int rbq_version(void) { return RBQ_VERSION; }
int rbq_set_option(rbq_handle* h, const char* name, const char* value) {
    if (!h || !name) return RBQ_E_NULL;
    h->err[0] = 0;
    const int rc = rbq_apply_option(h, name, value);
    if (rc == RBQ_E_NULL) return fail(h, RBQ_E_SIZE, "rbq_set_option: unknown option %s", name);
    if (rc) return fail(h, rc, "rbq_set_option: bad value for %s", name);
    return RBQ_OK;
}

int rbq_create(int device, rbq_handle** out) {
    if (!out) return RBQ_E_NULL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return RBQ_E_NODEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RBQ_E_NODEVICE;
    if (prop.major != 10) return RBQ_E_NODEVICE;   // kernels are built for sm_100a only
    rbq_handle* h = new (std::nothrow) rbq_handle();
    if (!h) return RBQ_E_SIZE;
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->centre_fq = 1.4e-2; h->centre_da = 0.0;   // FQ (spin.jl:148), no amplitude offset
    // tuning switches are read from the environment here and nowhere else (per handle, immutable unless rbq_set_option)
    for (const char* name : kOptionNames) {
        if (rbq_apply_option(h, name, nullptr) != RBQ_OK) { delete h; return RBQ_E_SIZE; }
        const char* e = getenv(name);
        if (e && *e) rbq_apply_option(h, name, e);   // a malformed value keeps the default
    }
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return RBQ_E_NODEVICE; }
    bool ok = true;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaStreamCreateWithFlags(&h->copy_stream[i], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    if (!ok) {   // release whatever was created before the failure
        for (int i = 0; i < 2; ++i) if (h->copy_stream[i]) cudaStreamDestroy(h->copy_stream[i]);
        for (int i = 0; i < 8; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
        cudaGetLastError();
        delete h;
        return RBQ_E_NODEVICE;
    }
    *out = h;
    return RBQ_OK;
}
int rbq_destroy(rbq_handle* h) {
    if (!h) return RBQ_E_NULL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    ctab_forget(h);
    for (int i = 0; i < RBQ_NBUF; ++i) if (h->dbuf[i]) cudaFree(h->dbuf[i]);
    free(h->idx_host);
    for (int i = 0; i < 2; ++i) cudaStreamDestroy(h->copy_stream[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev[i]);
    delete h;
    return RBQ_OK;
}
const char* rbq_last_error(const rbq_handle* h) { return h ? h->err : "null handle"; }
int rbq_set_stream(rbq_handle* h, void* cuda_stream) {
    if (!h) return RBQ_E_NULL;
    h->stream = (cudaStream_t)cuda_stream;
    return RBQ_OK;
}
int rbq_set_sweep_centre(rbq_handle* h, double fq0, double da0) {
    ENTER(h);
    if (!isfinite(fq0) || !isfinite(da0)) return fail(h, RBQ_E_SIZE, "rbq_set_sweep_centre: fq0=%g da0=%g", fq0, da0);
    h->centre_fq = fq0; h->centre_da = da0;
    return RBQ_OK;
}
int rbq_synchronize(rbq_handle* h) {
    ENTER(h);
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int64_t rbq_launch_count(const rbq_handle* h) { return h ? h->launches : -1; }
int rbq_host_alloc(rbq_handle* h, int64_t bytes, void** out) {
    ENTER(h);
    if (!out) return fail(h, RBQ_E_NULL, "rbq_host_alloc: null pointer");
    *out = nullptr;
    if (bytes <= 0) return fail(h, RBQ_E_SIZE, "rbq_host_alloc: bytes=%lld", (long long)bytes);
    CK(cudaMallocHost(out, (size_t)bytes));
    return RBQ_OK;
}
int rbq_host_register(rbq_handle* h, void* ptr, int64_t bytes) {
    ENTER(h);
    if (!ptr) return fail(h, RBQ_E_NULL, "rbq_host_register: null pointer");
    if (bytes <= 0) return fail(h, RBQ_E_SIZE, "rbq_host_register: bytes=%lld", (long long)bytes);
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RBQ_OK;
}
int rbq_host_unregister(rbq_handle* h, void* ptr) {
    ENTER(h);
    if (ptr) CK(cudaHostUnregister(ptr));
    return RBQ_OK;
}
int rbq_host_free(rbq_handle* h, void* ptr) {
    // page-locked memory is process wide, not tied to a handle: h may be NULL (the owner of the buffer outlived its handle)
    if (h) { ENTER(h); }
    if (ptr) {
        cudaError_t e = cudaFreeHost(ptr);
        if (e != cudaSuccess) return h ? fail(h, (int)e, "cudaFreeHost: %s", cudaGetErrorString(e)) : (int)e;
    }
    return RBQ_OK;
}

int rbq_state_dim(const rbq_model_params* p) { return p ? model_state_dim(*p) : RBQ_E_NULL; }
int rbq_control_dim(const rbq_model_params* p) { return p ? model_control_dim(*p) : RBQ_E_NULL; }

// ---- trajectory level ----------------------------------------------------------------------------------------
int rbq_rollout_dev(rbq_handle* h, const rbq_model_params* p, int64_t B, int64_t N, const double* d_x0,
                    const double* d_U, const double* d_dt, double* d_X) {
    ENTER(h);
    if (!p || !d_x0 || !d_U || !d_dt || !d_X) return fail(h, RBQ_E_NULL, "rbq_rollout_dev: null pointer");
    if (B < 0 || N < 1) return fail(h, RBQ_E_SIZE, "rbq_rollout_dev: B=%lld N=%lld", (long long)B, (long long)N);
    if (B == 0) return RBQ_OK;
    CKL(launch_rollout(h, *p, B, N, 0.0, d_x0, d_U, d_dt, d_X));
    return RBQ_OK;
}
int rbq_jacobians_dev(rbq_handle* h, const rbq_model_params* p, int64_t K, const double* d_X, const double* d_U,
                      const double* d_t, const double* d_dt, double* d_AB) {
    ENTER(h);
    if (!p || !d_X || !d_U || !d_dt || !d_AB) return fail(h, RBQ_E_NULL, "rbq_jacobians_dev: null pointer");
    if (K < 0) return fail(h, RBQ_E_SIZE, "rbq_jacobians_dev: K=%lld", (long long)K);
    if (K == 0) return RBQ_OK;
    CKL(launch_jacobians(h, *p, K, d_X, d_U, d_t, d_dt, d_AB));
    return RBQ_OK;
}
int rbq_rollout(rbq_handle* h, const rbq_model_params* p, int64_t B, int64_t N, const double* x0, const double* U,
                const double* dt, double* X) {
    ENTER(h);
    if (!p || !x0 || !X || (N > 1 && (!U || !dt))) return fail(h, RBQ_E_NULL, "rbq_rollout: null pointer");
    const int n = model_state_dim(*p), m = model_control_dim(*p);
    if (n < 0 || m < 0) return fail(h, RBQ_E_MODEL, "rbq_rollout: bad model parameters");
    if (B < 0 || N < 1) return fail(h, RBQ_E_SIZE, "rbq_rollout: B=%lld N=%lld", (long long)B, (long long)N);
    if (B == 0) return RBQ_OK;
    void *dx0, *dU, *ddt, *dX;
    const size_t bx0 = sizeof(double) * B * n, bU = sizeof(double) * B * (N - 1) * m, bdt = sizeof(double) * (N - 1),
                 bX = sizeof(double) * B * N * n;
    CKL(dev_buf(h, B_IN0, bx0, &dx0)); CKL(dev_buf(h, B_IN1, bU, &dU)); CKL(dev_buf(h, B_IN2, bdt, &ddt));
    CKL(dev_buf(h, B_OUT0, bX, &dX));
    CK(cudaMemcpyAsync(dx0, x0, bx0, cudaMemcpyHostToDevice, h->stream));
    if (N > 1) {
        CK(cudaMemcpyAsync(dU, U, bU, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(ddt, dt, bdt, cudaMemcpyHostToDevice, h->stream));
    }
    CKL(launch_rollout(h, *p, B, N, 0.0, (double*)dx0, (double*)dU, (double*)ddt, (double*)dX));
    CK(cudaMemcpyAsync(X, dX, bX, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_rollout_closedloop(rbq_handle* h, const rbq_model_params* p, int64_t N, const double* Xref, const double* Uref,
                           const double* K, const double* d, const double* dt, int64_t nalpha, const double* alphas,
                           double* X, double* Uout) {
    ENTER(h);
    if (!p || !Xref || !Uref || !K || !d || !dt || !alphas || !X || !Uout) return fail(h, RBQ_E_NULL, "rbq_rollout_closedloop: null pointer");
    const int n = model_state_dim(*p), m = model_control_dim(*p);
    if (n < 0 || m < 0) return fail(h, RBQ_E_MODEL, "rbq_rollout_closedloop: bad model parameters");
    if (N < 2 || nalpha < 0) return fail(h, RBQ_E_SIZE, "rbq_rollout_closedloop: N=%lld nalpha=%lld", (long long)N, (long long)nalpha);
    if (nalpha == 0) return RBQ_OK;
    void *dXr, *dUr, *dK, *dd, *ddt, *dal, *dX, *dUo;
    const size_t bXr = sizeof(double) * N * n, bUr = sizeof(double) * (N - 1) * m, bK = sizeof(double) * (N - 1) * m * n,
                 bd = bUr, bdt = sizeof(double) * (N - 1), bal = sizeof(double) * nalpha, bX = bXr * nalpha, bUo = bUr * nalpha;
    CKL(dev_buf(h, B_IN0, bXr, &dXr)); CKL(dev_buf(h, B_IN1, bUr, &dUr)); CKL(dev_buf(h, B_IN2, bdt, &ddt));
    CKL(dev_buf(h, B_IN3, bK, &dK)); CKL(dev_buf(h, B_IN4, bd, &dd)); CKL(dev_buf(h, B_IN5, bal, &dal));
    CKL(dev_buf(h, B_OUT0, bX, &dX)); CKL(dev_buf(h, B_OUT1, bUo, &dUo));
    CK(cudaMemcpyAsync(dXr, Xref, bXr, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dUr, Uref, bUr, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dK, K, bK, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dd, d, bd, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(ddt, dt, bdt, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dal, alphas, bal, cudaMemcpyHostToDevice, h->stream));
    CKL(launch_closedloop(h, *p, N, (double*)dXr, (double*)dUr, (double*)dK, (double*)dd, (double*)ddt, nalpha, (double*)dal,
                          (double*)dX, (double*)dUo));
    CK(cudaMemcpyAsync(X, dX, bX, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(Uout, dUo, bUo, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_rollout_closedloop_dev(rbq_handle* h, const rbq_model_params* p, int64_t N, const double* d_Xref, const double* d_Uref,
                               const double* K, const double* d, const double* dt, int64_t nalpha, const double* alphas,
                               double* X, double* Uout) {
    ENTER(h);
    if (!p || !d_Xref || !d_Uref || !K || !d || !dt || !alphas || !X || !Uout) return fail(h, RBQ_E_NULL, "rbq_rollout_closedloop_dev: null pointer");
    const int n = model_state_dim(*p), m = model_control_dim(*p);
    if (n < 0 || m < 0) return fail(h, RBQ_E_MODEL, "rbq_rollout_closedloop_dev: bad model parameters");
    if (N < 2 || nalpha < 0) return fail(h, RBQ_E_SIZE, "rbq_rollout_closedloop_dev: N=%lld nalpha=%lld", (long long)N, (long long)nalpha);
    if (nalpha == 0) return RBQ_OK;
    CKL(launch_closedloop(h, *p, N, d_Xref, d_Uref, K, d, dt, nalpha, alphas, X, Uout));
    return RBQ_OK;
}
int rbq_jacobians(rbq_handle* h, const rbq_model_params* p, int64_t K, const double* X, const double* U, const double* t,
                  const double* dt, double* AB) {
    ENTER(h);
    if (!p || !X || !U || !dt || !AB) return fail(h, RBQ_E_NULL, "rbq_jacobians: null pointer");
    const int n = model_state_dim(*p), m = model_control_dim(*p);
    if (n < 0 || m < 0) return fail(h, RBQ_E_MODEL, "rbq_jacobians: bad model parameters");
    if (K < 0) return fail(h, RBQ_E_SIZE, "rbq_jacobians: K=%lld", (long long)K);
    if (K == 0) return RBQ_OK;
    void *dX, *dU, *ddt, *dAB, *dtk = nullptr;
    const size_t bX = sizeof(double) * K * n, bU = sizeof(double) * K * m, bdt = sizeof(double) * K,
                 bAB = sizeof(double) * K * n * (n + m);
    CKL(dev_buf(h, B_IN0, bX, &dX)); CKL(dev_buf(h, B_IN1, bU, &dU)); CKL(dev_buf(h, B_IN2, bdt, &ddt));
    CKL(dev_buf(h, B_OUT0, bAB, &dAB));
    CK(cudaMemcpyAsync(dX, X, bX, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dU, U, bU, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(ddt, dt, bdt, cudaMemcpyHostToDevice, h->stream));
    if (t) {
        CKL(dev_buf(h, B_IN3, bdt, &dtk));
        CK(cudaMemcpyAsync(dtk, t, bdt, cudaMemcpyHostToDevice, h->stream));
    }
    CKL(launch_jacobians(h, *p, K, (double*)dX, (double*)dU, (double*)dtk, (double*)ddt, (double*)dAB));
    CK(cudaMemcpyAsync(AB, dAB, bAB, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
// per-knot forms: a rollout of two knots / a one-knot Jacobian batch
int rbq_discrete_dynamics(rbq_handle* h, const rbq_model_params* p, const double* x, const double* u, double t, double dt,
                          double* xnext) {
    ENTER(h);   // `t` matters for spin25 only (Int(div(time, dt)) selects the filter start-up branch, spin25.jl:196-214)
    if (!p || !x || !u || !xnext) return fail(h, RBQ_E_NULL, "rbq_discrete_dynamics: null pointer");
    const int n = model_state_dim(*p), m = model_control_dim(*p);
    if (n < 0 || m < 0) return fail(h, RBQ_E_MODEL, "rbq_discrete_dynamics: bad model parameters");
    void *dx, *du, *ddt, *dX;
    CKL(dev_buf(h, B_IN0, sizeof(double) * n, &dx)); CKL(dev_buf(h, B_IN1, sizeof(double) * m, &du));
    CKL(dev_buf(h, B_IN2, sizeof(double), &ddt)); CKL(dev_buf(h, B_OUT0, sizeof(double) * 2 * n, &dX));
    CK(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(du, u, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(ddt, &dt, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CKL(launch_rollout(h, *p, 1, 2, t, (double*)dx, (double*)du, (double*)ddt, (double*)dX));
    CK(cudaMemcpyAsync(xnext, (double*)dX + n, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_discrete_jacobian(rbq_handle* h, const rbq_model_params* p, const double* x, const double* u, double t, double dt,
                          double* AB) {
    return rbq_jacobians(h, p, 1, x, u, &t, &dt, AB);
}

// ---- cost-side primitives ------------------------------------------------------------------------------------------
int rbq_gate_error_expansion_dev(rbq_handle* h, int64_t K, const double* d_s1, int64_t s1_stride, const double* d_s2,
                                 int64_t s2_stride, double* d_err, double* d_grad, double* d_hess) {
    ENTER(h);
    if (K < 0 || s1_stride < 4 || (s2_stride != 0 && s2_stride < 4)) return fail(h, RBQ_E_SIZE, "rbq_gate_error_expansion_dev: K=%lld strides %lld/%lld", (long long)K, (long long)s1_stride, (long long)s2_stride);
    if (K == 0) return RBQ_OK;
    if (!d_s1 || !d_s2) return fail(h, RBQ_E_NULL, "rbq_gate_error_expansion_dev: null pointer");
    CKL(launch_gate_error(h, K, d_s1, s1_stride, d_s2, s2_stride, d_err, d_grad, d_hess));
    return RBQ_OK;
}
int rbq_gate_error_expansion(rbq_handle* h, int64_t K, const double* s1, int64_t s1_stride, const double* s2, int64_t s2_stride,
                             double* err, double* grad, double* hess) {
    ENTER(h);
    if (K < 0 || s1_stride < 4 || (s2_stride != 0 && s2_stride < 4)) return fail(h, RBQ_E_SIZE, "rbq_gate_error_expansion: K=%lld strides %lld/%lld", (long long)K, (long long)s1_stride, (long long)s2_stride);
    if (K == 0) return RBQ_OK;
    if (!s1 || !s2) return fail(h, RBQ_E_NULL, "rbq_gate_error_expansion: null pointer");
    const size_t b1 = sizeof(double) * ((size_t)(K - 1) * s1_stride + 4), b2 = sizeof(double) * ((size_t)(K - 1) * s2_stride + 4);
    void *d1, *d2, *de = nullptr, *dg = nullptr, *dh = nullptr;
    CKL(dev_buf(h, B_IN0, b1, &d1)); CKL(dev_buf(h, B_IN1, b2, &d2));
    if (err) CKL(dev_buf(h, B_OUT0, sizeof(double) * K, &de));
    if (grad) CKL(dev_buf(h, B_OUT1, sizeof(double) * 4 * K, &dg));
    if (hess) CKL(dev_buf(h, B_IN4, sizeof(double) * 16 * K, &dh));
    CK(cudaMemcpyAsync(d1, s1, b1, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d2, s2, b2, cudaMemcpyHostToDevice, h->stream));
    CKL(launch_gate_error(h, K, (double*)d1, s1_stride, (double*)d2, s2_stride, (double*)de, (double*)dg, (double*)dh));
    if (err) CK(cudaMemcpyAsync(err, de, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
    if (grad) CK(cudaMemcpyAsync(grad, dg, sizeof(double) * 4 * K, cudaMemcpyDeviceToHost, h->stream));
    if (hess) CK(cudaMemcpyAsync(hess, dh, sizeof(double) * 16 * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

int rbq_sampling_cost_expansion(rbq_handle* h, int64_t K, const double* X, const double* U, const double* Qdiag,
                                const double* Rdiag, const double* xf, const double* targets, const double* q_ss, double* J,
                                double* gx, double* gu) {
    ENTER(h);
    if (K < 0) return fail(h, RBQ_E_SIZE, "rbq_sampling_cost_expansion: K=%lld", (long long)K);
    if (K == 0) return RBQ_OK;
    if (!X || !Qdiag || !xf || !targets || !q_ss || (U && !Rdiag)) return fail(h, RBQ_E_NULL, "rbq_sampling_cost_expansion: null pointer");
    void *dX, *dU = nullptr, *dJ = nullptr, *dgx = nullptr, *dgu = nullptr;
    CKL(dev_buf(h, B_IN0, sizeof(double) * 43 * K, &dX));
    CK(cudaMemcpyAsync(dX, X, sizeof(double) * 43 * K, cudaMemcpyHostToDevice, h->stream));
    if (U) {
        CKL(dev_buf(h, B_IN1, sizeof(double) * K, &dU));
        CK(cudaMemcpyAsync(dU, U, sizeof(double) * K, cudaMemcpyHostToDevice, h->stream));
    }
    if (J) CKL(dev_buf(h, B_OUT0, sizeof(double) * K, &dJ));
    if (gx) CKL(dev_buf(h, B_OUT1, sizeof(double) * 43 * K, &dgx));
    if (gu && U) CKL(dev_buf(h, B_IN4, sizeof(double) * K, &dgu));
    CKL(launch_sampling_cost(h, K, (double*)dX, (double*)dU, Qdiag, U ? Rdiag[0] : 0.0, xf, targets, q_ss, (double*)dJ, (double*)dgx,
                             (double*)dgu));
    if (J) CK(cudaMemcpyAsync(J, dJ, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
    if (gx) CK(cudaMemcpyAsync(gx, dgx, sizeof(double) * 43 * K, cudaMemcpyDeviceToHost, h->stream));
    if (gu && U) CK(cudaMemcpyAsync(gu, dgu, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- Monte-Carlo sweeps ----------------------------------------------------------------------------------------
static int check_mc(rbq_handle* h, const char* who, const void* controls, int64_t nk, double dt, int64_t gate_count, int64_t S,
                    int algo) {
    if (!controls) return fail(h, RBQ_E_NULL, "%s: controls is NULL", who);
    if (nk < 1 || nk > ((int64_t)1 << 40) || gate_count < 0 || S < 0) return fail(h, RBQ_E_SIZE, "%s: nknots=%lld gate_count=%lld S=%lld", who, (long long)nk, (long long)gate_count, (long long)S);
    if (!(dt == dt)) return fail(h, RBQ_E_SIZE, "%s: dt is NaN", who);
    if (algo != RBQ_ALGO_SU2 && algo != RBQ_ALGO_PADE) return fail(h, RBQ_E_ALGO, "%s: unknown algo %d", who, algo);
    return 0;
}
// stage the pulse, return device pointers to the staged copy and to max|a|
static int stage_pulse(rbq_handle* h, const double* d_controls, int64_t nk, const double** staged, const double** amax) {
    void *pb = nullptr, *sc = nullptr;
    CKL(dev_buf(h, B_PULSE, sizeof(double) * (nk + 2), &pb));
    CKL(dev_buf(h, B_SCALARS, 256, &sc));
    CKL(launch_pulse_prep(h, d_controls, nk, (double*)sc));
    *staged = (const double*)pb; *amax = (const double*)sc;
    return 0;
}
int rbq_stats_reset_dev(rbq_handle* h, rbq_stats* d_stats) {
    ENTER(h);
    if (!d_stats) return fail(h, RBQ_E_NULL, "rbq_stats_reset_dev: null pointer");
    CKL(launch_stats_reset(h, d_stats));
    return RBQ_OK;
}
int rbq_stats_merge(rbq_stats* a, const rbq_stats* b) {
    if (!a || !b) return RBQ_E_NULL;
    a->count += b->count; a->nonfinite += b->nonfinite; a->sum += b->sum; a->sumsq += b->sumsq;
    a->min = fmin(a->min, b->min); a->max = fmax(a->max, b->max);
    for (int i = 0; i < RBQ_HIST_BINS; ++i) a->hist[i] += b->hist[i];
    return RBQ_OK;
}

// build the knot table (expanded around the sweep centre) and fill everything of McStaticArgs that does not depend on
// the sample range
static int mc_static_prepare(rbq_handle* h, const double* d_controls, int64_t nk, double dt, int gate, int64_t gate_count,
                             int64_t partial_blocks, rbq_stats* d_stats, double f0, double a0, int64_t launch_samples,
                             McStaticArgs* a) {
    memset(a, 0, sizeof(*a));
    if (!make_gate_tables(gate, &a->gate, nullptr)) return fail(h, RBQ_E_GATE, "unknown gate %d", gate);
    void *tb = nullptr, *sc = nullptr;
    CKL(dev_buf(h, B_PULSE, sizeof(double) * 8 * (size_t)nk + 32, &tb));
    CKL(dev_buf(h, B_SCALARS, 256, &sc));
    CKL(launch_static_table(h, d_controls, nk, dt, f0, a0, (double*)tb, (double*)sc, (double*)tb + 4 * (size_t)nk));
    a->table = (const double*)tb; a->scal = (const double*)sc;
    a->table_m = (const double*)tb + 4 * (size_t)nk; a->table_gen = ++h->table_gen;
    {   // quaternion scratch of the constant-memory form, sized for the largest launch of this call
        void* qb = nullptr;
        CKL(dev_buf(h, B_QUAT, sizeof(double) * 4 * (size_t)(launch_samples > 0 ? launch_samples : 1), &qb));
        a->quat = (double*)qb;
    }
    a->no_table_class = h->opt.no_table_class;
    a->nk = nk; a->dt = dt; a->gate_count = gate_count;
    if (d_stats) {
        void* pp;
        CKL(dev_buf(h, B_PARTIALS, sizeof(BlockPartial) * (size_t)(partial_blocks > 0 ? partial_blocks : 1), &pp));
        a->partials = (BlockPartial*)pp;
        a->hist = (unsigned long long*)d_stats->hist;
    }
    return 0;
}
static int mc_static_core(rbq_handle* h, const double* d_controls, int64_t nk, double dt, int gate, int64_t gate_count,
                          int64_t first, int64_t S, const double* d_fq, const double* d_da, const double* d_psi0, int philox,
                          uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip, double da_sigma, int algo,
                          double* d_fid, rbq_stats* d_stats, int host_io = 0) {
    // host_io: the sample / fidelity pointers are device addresses of page-locked HOST memory (the zero-copy entry points):
    // such a sweep keeps the one-kernel form, whose blocks hide their PCIe traffic behind their neighbours' arithmetic
    McStaticArgs a;
    if (S == 0) { GateIso gi; return make_gate_tables(gate, &gi, nullptr) ? RBQ_OK : fail(h, RBQ_E_GATE, "unknown gate %d", gate); }
    // Philox sweeps are centred on their own (fq0, 0), explicit samples on the handle's centre (rbq_set_sweep_centre)
    CKL(mc_static_prepare(h, d_controls, nk, dt, gate, gate_count, mc_static_blocks(h, S, algo), d_stats,
                          philox ? fq0 : h->centre_fq, philox ? 0.0 : h->centre_da, S, &a));
    a.first = first; a.S = S;
    a.fq = d_fq; a.da = d_da; a.psi0 = d_psi0;
    a.philox = philox; a.seed = seed; a.fq0 = fq0; a.fq_rel_sigma = fq_rel_sigma; a.fq_rel_clip = fq_rel_clip; a.da_sigma = da_sigma;
    a.fid = d_fid;
    a.host_io = host_io;
    int nblocks = 0;
    CKL(launch_mc_static(h, a, algo, &nblocks));
    if (d_stats) CKL(launch_stats_finalize(h, a.partials, nblocks, d_stats));
    return RBQ_OK;
}
int rbq_mc_static_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                      int64_t S, const double* d_fq, const double* d_da, const double* d_psi0, int algo, double* d_fid,
                      rbq_stats* d_stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_static_dev", d_controls, nknots, dt, gate_count, S, algo));
    if (S > 0 && (!d_fq || !d_psi0)) return fail(h, RBQ_E_NULL, "rbq_mc_static_dev: fq/psi0 is NULL");
    return mc_static_core(h, d_controls, nknots, dt, gate, gate_count, 0, S, d_fq, d_da, d_psi0, 0, 0, 0, 0, 0, 0, algo, d_fid, d_stats);
}
int rbq_mc_static_philox_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                             int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                             double da_sigma, int algo, double* d_fid, rbq_stats* d_stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_static_philox_dev", d_controls, nknots, dt, gate_count, S, algo));
    if (first < 0) return fail(h, RBQ_E_SIZE, "rbq_mc_static_philox_dev: first=%lld", (long long)first);
    return mc_static_core(h, d_controls, nknots, dt, gate, gate_count, first, S, nullptr, nullptr, nullptr, 1, seed, fq0,
                          fq_rel_sigma, fq_rel_clip, da_sigma, algo, d_fid, d_stats);
}

// host-pointer sweep: samples are streamed through the device in chunks so that the H2D copy of chunk c+1, the kernel
// of chunk c and the D2H copy of chunk c-1 overlap; the pulse is staged once and the statistics finalised once.
// chunk c covers [mc_chunk_begin(c), mc_chunk_begin(c+1)): whole waves of resident blocks, so chunking adds no partially
// filled wave and the upload of wave c+1 hides behind the kernel of wave c
static inline int64_t mc_chunk_begin(const rbq_handle* h, int64_t c, int64_t wave, int64_t S) {
    const int64_t b = c == 0 ? 0 : wave * (h->opt.chunk_first + (int64_t)h->opt.chunk_rest * (c - 1));
    return b < S ? b : S;
}
// page-locked host memory the device can address directly (cudaMallocHost / cudaHostRegister / torch pin_memory)
static bool host_mapped(const void* p, void** dptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *dptr = at.devicePointer;
    return true;
}
int rbq_mc_static(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count, int64_t S,
                  const double* fq, const double* da, const double* psi0, int algo, double* fid, rbq_stats* stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_static", controls, nknots, dt, gate_count, S, algo));
    if (S > 0 && (!fq || !psi0)) return fail(h, RBQ_E_NULL, "rbq_mc_static: fq/psi0 is NULL");
    GateIso gi;
    if (!make_gate_tables(gate, &gi, nullptr)) return fail(h, RBQ_E_GATE, "rbq_mc_static: unknown gate %d", gate);
    // Zero-copy form: when every sample buffer is page-locked, ONE kernel reads the samples (48 B each, once, coalesced)
    // and writes the fidelities straight over PCIe; block start-up staggers the transfers behind the arithmetic of the
    // other resident blocks, so there are no chunk boundaries, launch gaps or staging copies.  Worth it while the
    // kernel is compute bound, i.e. few recorded gates per sample.  RBQ_MC_ZEROCOPY=0 forces the streamed form.
    {
        void *mfq = nullptr, *mda = nullptr, *mpsi = nullptr, *mfid = nullptr;
        const bool want = h->opt.mc_zerocopy && algo == RBQ_ALGO_SU2 && S > 0 && gate_count <= 3;
        if (want && host_mapped(fq, &mfq) && host_mapped(psi0, &mpsi) && (!da || host_mapped(da, &mda)) &&
            (!fid || host_mapped(fid, &mfid))) {
            void *dc0, *dst0 = nullptr;
            CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc0));
            if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst0)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst0)); }
            CK(cudaMemcpyAsync(dc0, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
            int rc = mc_static_core(h, (double*)dc0, nknots, dt, gate, gate_count, 0, S, (double*)mfq, (double*)mda, (double*)mpsi,
                                    0, 0, 0, 0, 0, 0, algo, (double*)mfid, (rbq_stats*)dst0, 1);
            if (rc) return rc;
            if (stats) CK(cudaMemcpyAsync(stats, dst0, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            return RBQ_OK;
        }
    }
    void *dc, *dfq, *dda, *dpsi, *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc));
    CKL(dev_buf(h, B_IN1, sizeof(double) * S, &dfq));
    CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda));
    CKL(dev_buf(h, B_IN3, sizeof(double) * S * 4, &dpsi));
    if (fid) CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, &dfid));
    if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst)); }
    CK(cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
    if (S > 0) {
        const int64_t wave = mc_static_wave_samples(h, algo);
        int64_t nchunks = 0;
        while (mc_chunk_begin(h, nchunks, wave, S) < S) ++nchunks;
        int64_t total_blocks = 0, largest = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t s0 = mc_chunk_begin(h, c, wave, S), cnt = mc_chunk_begin(h, c + 1, wave, S) - s0;
            total_blocks += mc_static_blocks(h, cnt, algo);
            largest = cnt > largest ? cnt : largest;
        }
        McStaticArgs a;
        CKL(mc_static_prepare(h, (double*)dc, nknots, dt, gate, gate_count, total_blocks, (rbq_stats*)dst, h->centre_fq,
                              h->centre_da, largest, &a));
        BlockPartial* partials0 = a.partials;
        // event choreography: copy_stream[0] uploads, h->stream computes, copy_stream[1] downloads
        cudaStream_t up = h->copy_stream[0], down = h->copy_stream[1];
        CK(cudaEventRecord(h->ev[0], h->stream));
        CK(cudaStreamWaitEvent(up, h->ev[0], 0));   // uploads start after earlier work on the buffers has drained
        int64_t block_off = 0;
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t s0 = mc_chunk_begin(h, c, wave, S), cnt = mc_chunk_begin(h, c + 1, wave, S) - s0;
            CK(cudaMemcpyAsync((double*)dfq + s0, fq + s0, sizeof(double) * cnt, cudaMemcpyHostToDevice, up));
            if (da) CK(cudaMemcpyAsync((double*)dda + s0, da + s0, sizeof(double) * cnt, cudaMemcpyHostToDevice, up));
            CK(cudaMemcpyAsync((double*)dpsi + 4 * s0, psi0 + 4 * s0, sizeof(double) * 4 * cnt, cudaMemcpyHostToDevice, up));
            CK(cudaEventRecord(h->ev[1 + (c & 1)], up));
            CK(cudaStreamWaitEvent(h->stream, h->ev[1 + (c & 1)], 0));
            a.first = 0; a.S = cnt;
            a.fq = (double*)dfq + s0; a.da = da ? (double*)dda + s0 : nullptr; a.psi0 = (double*)dpsi + 4 * s0;
            a.fid = fid ? (double*)dfid + s0 * G1 : nullptr;
            if (partials0) a.partials = partials0 + block_off;
            int nblocks = 0;
            CKL(launch_mc_static(h, a, algo, &nblocks));
            block_off += nblocks;
            if (fid) {
                CK(cudaEventRecord(h->ev[3 + (c & 1)], h->stream));
                CK(cudaStreamWaitEvent(down, h->ev[3 + (c & 1)], 0));
                CK(cudaMemcpyAsync(fid + s0 * G1, (double*)dfid + s0 * G1, sizeof(double) * cnt * G1, cudaMemcpyDeviceToHost, down));
            }
        }
        if (stats) CKL(launch_stats_finalize(h, partials0, (int)block_off, (rbq_stats*)dst));
        CK(cudaStreamSynchronize(down));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_mc_static_philox(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                         int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                         double da_sigma, int algo, double* fid, rbq_stats* stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_static_philox", controls, nknots, dt, gate_count, S, algo));
    void *dc, *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc));
    // page-locked fid (rbq_host_alloc / rbq_host_register / torch pin_memory) and few recorded gates: the kernel writes the
    // fidelities straight to host memory while it computes (no staging buffer, no trailing copy), as rbq_mc_static does
    void* mfid = nullptr;
    const bool direct = fid && S > 0 && h->opt.mc_zerocopy && algo == RBQ_ALGO_SU2 && gate_count <= 3 && host_mapped(fid, &mfid);
    if (fid && !direct) CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, &dfid));
    if (direct) dfid = mfid;
    if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst)); }
    CK(cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
    if (first < 0) return fail(h, RBQ_E_SIZE, "rbq_mc_static_philox: first=%lld", (long long)first);
    int rc = mc_static_core(h, (double*)dc, nknots, dt, gate, gate_count, first, S, nullptr, nullptr, nullptr, 1, seed, fq0,
                            fq_rel_sigma, fq_rel_clip, da_sigma, algo, (double*)dfid, (rbq_stats*)dst, direct ? 1 : 0);
    if (rc) return rc;
    if (fid && S > 0 && !direct) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
// ---- single-process multi-device group ------------------------------------------------------------------------------------
struct rbq_group {
    int n;
    rbq_handle* h[64];
    cudaStream_t own_stream[64];     // members get their own non-blocking stream so that their work overlaps
    char err[512];
};
int rbq_group_create(const int* devices, int n, rbq_group** out) {
    if (!devices || !out) return RBQ_E_NULL;
    *out = nullptr;
    if (n < 1 || n > 64) return RBQ_E_SIZE;
    rbq_group* g = new (std::nothrow) rbq_group();
    if (!g) return RBQ_E_SIZE;
    memset(g, 0, sizeof(*g));
    for (int i = 0; i < n; ++i) {
        int rc = rbq_create(devices[i], &g->h[i]);
        if (rc == RBQ_OK && (cudaSetDevice(devices[i]) != cudaSuccess ||
                             cudaStreamCreateWithFlags(&g->own_stream[i], cudaStreamNonBlocking) != cudaSuccess)) rc = RBQ_E_NODEVICE;
        if (rc != RBQ_OK) {
            g->n = i + (g->h[i] ? 1 : 0);
            rbq_group_destroy(g);
            return rc;
        }
        g->h[i]->stream = g->own_stream[i];
        g->n = i + 1;
    }
    *out = g;
    return RBQ_OK;
}
int rbq_group_destroy(rbq_group* g) {
    if (!g) return RBQ_E_NULL;
    for (int i = 0; i < g->n; ++i) {
        if (!g->h[i]) continue;
        const int dev = g->h[i]->device;
        rbq_destroy(g->h[i]);
        if (g->own_stream[i]) { cudaSetDevice(dev); cudaStreamDestroy(g->own_stream[i]); }
    }
    delete g;
    return RBQ_OK;
}
int rbq_group_size(const rbq_group* g) { return g ? g->n : RBQ_E_NULL; }
rbq_handle* rbq_group_handle(rbq_group* g, int member) { return (g && member >= 0 && member < g->n) ? g->h[member] : nullptr; }
const char* rbq_group_last_error(const rbq_group* g) { return g ? g->err : "null group"; }
int rbq_group_mc_static_philox(rbq_group* g, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                               int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma, double fq_rel_clip,
                               double da_sigma, int algo, double* fid, rbq_stats* stats) {
    if (!g) return RBQ_E_NULL;
    g->err[0] = 0;
    if (S < 0 || first < 0) { snprintf(g->err, sizeof(g->err), "rbq_group_mc_static_philox: first=%lld S=%lld", (long long)first, (long long)S); return RBQ_E_SIZE; }
    const int64_t G1 = gate_count + 1;
    const int64_t base = S / g->n, rem = S % g->n;
    void* dst[64] = {nullptr};
    void* dfid[64] = {nullptr};
    int64_t off[65];
    off[0] = 0;
    for (int i = 0; i < g->n; ++i) off[i + 1] = off[i] + base + (i < rem ? 1 : 0);
    auto fail_with = [&](int i, int rc) { snprintf(g->err, sizeof(g->err), "member %d (device %d): %s", i, g->h[i]->device, g->h[i]->err); return rc; };
    // 1. enqueue every member's share (uploads, kernels, downloads) without waiting
    for (int i = 0; i < g->n; ++i) {
        rbq_handle* h = g->h[i];
        const int64_t cnt = off[i + 1] - off[i];
        if (cudaSetDevice(h->device) != cudaSuccess) return fail_with(i, RBQ_E_NODEVICE);
        h->err[0] = 0;
        int rc = check_mc(h, "rbq_group_mc_static_philox", controls, nknots, dt, gate_count, cnt, algo);
        if (rc) return fail_with(i, rc);
        void* dc;
        if ((rc = dev_buf(h, B_IN0, sizeof(double) * nknots, &dc))) return fail_with(i, rc);
        if (fid && cnt > 0 && (rc = dev_buf(h, B_OUT0, sizeof(double) * cnt * G1, &dfid[i]))) return fail_with(i, rc);
        if ((rc = dev_buf(h, B_STATS, sizeof(rbq_stats), &dst[i]))) return fail_with(i, rc);
        if ((rc = rbq_stats_reset_dev(h, (rbq_stats*)dst[i]))) return fail_with(i, rc);
        if (cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return fail_with(i, (int)cudaGetLastError());
        rc = rbq_mc_static_philox_dev(h, (double*)dc, nknots, dt, gate, gate_count, first + off[i], cnt, seed, fq0, fq_rel_sigma, fq_rel_clip,
                                      da_sigma, algo, (double*)dfid[i], (rbq_stats*)dst[i]);
        if (rc) return fail_with(i, rc);
        if (fid && cnt > 0 &&
            cudaMemcpyAsync(fid + off[i] * G1, dfid[i], sizeof(double) * cnt * G1, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
            return fail_with(i, (int)cudaGetLastError());
    }
    // 2. wait for each, merge the statistics in member order
    rbq_stats total;
    memset(&total, 0, sizeof(total));
    total.min = INFINITY; total.max = -INFINITY;
    for (int i = 0; i < g->n; ++i) {
        rbq_handle* h = g->h[i];
        rbq_stats part;
        if (cudaSetDevice(h->device) != cudaSuccess) return fail_with(i, RBQ_E_NODEVICE);
        if (cudaMemcpyAsync(&part, dst[i], sizeof(part), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            snprintf(h->err, sizeof(h->err), "%s", cudaGetErrorString(cudaGetLastError()));
            return fail_with(i, 1);
        }
        rbq_stats_merge(&total, &part);
    }
    if (stats) *stats = total;
    return RBQ_OK;
}

int rbq_philox_samples(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double fq0, double fq_rel_sigma,
                       double fq_rel_clip, double da_sigma, double* fq, double* da, double* psi0) {
    ENTER(h);
    if (S < 0 || first < 0) return fail(h, RBQ_E_SIZE, "rbq_philox_samples: first=%lld S=%lld", (long long)first, (long long)S);
    if (S == 0) return RBQ_OK;
    void *dfq, *dda, *dpsi;
    CKL(dev_buf(h, B_IN1, sizeof(double) * S, &dfq)); CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda));
    CKL(dev_buf(h, B_IN3, sizeof(double) * 4 * S, &dpsi));
    CKL(launch_philox_samples(h, first, S, seed, fq0, fq_rel_sigma, fq_rel_clip, da_sigma, (double*)dfq, (double*)dda, (double*)dpsi));
    if (fq) CK(cudaMemcpyAsync(fq, dfq, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
    if (da) CK(cudaMemcpyAsync(da, dda, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
    if (psi0) CK(cudaMemcpyAsync(psi0, dpsi, sizeof(double) * 4 * S, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- time-dependent noise -----------------------------------------------------------------------------------------
// Noise-sample index of every (gate, knot): idx = floor(time * noise_dt_inv) with `time` accumulated by repeated
// `time += dt` exactly as integrate_prop_schroedda! does (spin.jl:472-481; 0-based, clamped to the last sample).  The
// sequence does not depend on the sample, so it is built once on the host, cached in the handle and streamed to the
// kernel next to the pulse.
static int noise_index_table(rbq_handle* h, int64_t nk, double dt, int64_t gates, double ndi, int64_t noise_len,
                             const int32_t** d_table, int64_t* stride) {
    const int64_t nkp = (nk + 3) & ~(int64_t)3;
    *stride = nkp;
    if (noise_len >= ((int64_t)1 << 31)) return fail(h, RBQ_E_SIZE, "noise_len %lld does not fit the index table", (long long)noise_len);
    const size_t count = (size_t)nkp * (size_t)(gates > 0 ? gates : 1);
    void* dtab;
    const bool dev_ok = h->dcap[B_IDX] >= count * sizeof(int32_t);
    CKL(dev_buf(h, B_IDX, count * sizeof(int32_t), &dtab));
    *d_table = (const int32_t*)dtab;
    if (dev_ok && h->idx_host && h->idx_nk == nk && h->idx_gates == gates && h->idx_len == noise_len && h->idx_dt == dt && h->idx_ndi == ndi)
        return 0;
    if (h->idx_cap < count) {
        free(h->idx_host);
        h->idx_host = (int32_t*)malloc(count * sizeof(int32_t));
        h->idx_cap = h->idx_host ? count : 0;
        if (!h->idx_host) return fail(h, RBQ_E_SIZE, "out of host memory for the noise index table");
    }
    volatile double time = 0.0;   // volatile: no fused or extended-precision shortcuts, the sum is the reference's
    for (int64_t g = 0; g < gates; ++g) {
        int32_t* rowp = h->idx_host + g * nkp;
        for (int64_t j = 0; j < nk; ++j) {
            const double prod = time * ndi;
            int64_t idx = (int64_t)floor(prod);
            if (!(idx < noise_len - 1)) idx = noise_len - 1;
            if (idx < 0) idx = 0;
            rowp[j] = (int32_t)idx;
            time = time + dt;
        }
        for (int64_t j = nk; j < nkp; ++j) rowp[j] = rowp[nk - 1];
    }
    CK(cudaStreamSynchronize(h->stream));   // an earlier launch may still be reading the device copy
    CK(cudaMemcpyAsync(dtab, h->idx_host, count * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // the host copy is pageable
    h->idx_nk = nk; h->idx_gates = gates; h->idx_len = noise_len; h->idx_dt = dt; h->idx_ndi = ndi;
    return 0;
}

static int mc_noise_core(rbq_handle* h, const double* d_controls, int64_t nk, double dt, int gate, int64_t gate_count,
                         int64_t first, int64_t S, double fq, const double* d_noise, int64_t noise_len, double noise_dt_inv,
                         const double* d_psi0, int philox, uint64_t seed, double namp, int algo, double* d_fid,
                         rbq_stats* d_stats) {
    McNoiseArgs a;
    memset(&a, 0, sizeof(a));
    if (!make_gate_tables(gate, &a.gate, nullptr)) return fail(h, RBQ_E_GATE, "unknown gate %d", gate);
    if (S == 0) return RBQ_OK;
    {   // knot table around (fq, 0): every sample of a noise sweep shares the qubit frequency
        void *tb = nullptr, *sc = nullptr;
        CKL(dev_buf(h, B_PULSE, sizeof(double) * 4 * (size_t)nk + 32, &tb));
        CKL(dev_buf(h, B_SCALARS, 256, &sc));
        CKL(launch_static_table(h, d_controls, nk, dt, fq, 0.0, (double*)tb, (double*)sc));
        a.table = (const double*)tb; a.scal = (const double*)sc;
        a.no_table_class = h->opt.no_table_class;
    }
    a.nk = nk; a.dt = dt; a.gate_count = gate_count; a.first = first; a.S = S; a.fq = fq;
    a.noise = d_noise; a.noise_len = noise_len; a.noise_dt_inv = noise_dt_inv; a.psi0 = d_psi0;
    a.philox = philox; a.seed = seed; a.namp = namp; a.fid = d_fid;
    CKL(noise_index_table(h, nk, dt, gate_count, noise_dt_inv, noise_len, &a.idx_table, &a.idx_stride));
    if (d_stats) {
        void* pp;
        CKL(dev_buf(h, B_PARTIALS, sizeof(BlockPartial) * mc_noise_blocks(h, S), &pp));
        a.partials = (BlockPartial*)pp;
        a.hist = (unsigned long long*)d_stats->hist;
    }
    int nblocks = 0;
    // Few samples, several gates: S chains of gate_count * nk dependent knots cannot fill the GPU, S * gate_count gate
    // propagators can (see mc_noise_gate_kernel).  Large sweeps keep the single-kernel form (no scratch, same throughput).
    bool gate_parallel = algo == RBQ_ALGO_SU2 && gate_count >= 2 && gate_count <= 65535 && S < (int64_t)h->sm_count * 4096 &&
                         (double)S * (double)gate_count * 32.0 <= 2.0e9;
    // Seed-in sweeps in the gate-parallel form need their noise rows in memory (the pink filter is sequential in time, a gate's
    // thread cannot start in the middle of it): they are materialised first (pink_white_kernel + pink_filter_kernel) whenever they
    // fit -- a quarter of the free device memory, at most 32 GB.  The single-kernel form keeps generating its noise inside the
    // sweep: materialising measured slower there (2^20 samples x 1 gate: 16.5 against 13.2 ms; 65 536 x 1: 1.21 against 1.02).
    bool materialise = false;
    if (philox && gate_parallel) {
        const double bytes = (double)S * (double)noise_len * 8.0;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
        const double have = (double)free_b + (double)h->dcap[B_NOISE];        // the scratch this handle already holds counts as free
        materialise = h->opt.noise_materialise != 0 && bytes <= 32.0e9 && bytes <= 0.25 * have;
        if (!materialise) gate_parallel = false;
    }
    if (h->opt.noise_gate_parallel == 0) { gate_parallel = false; materialise = false; }
    if (materialise) {
        void* dn;
        CKL(dev_buf(h, B_NOISE, sizeof(double) * (size_t)S * (size_t)noise_len, &dn));
        CKL(launch_pink_noise(h, first, S, seed, namp, noise_dt_inv, noise_len, (double*)dn));
        a.noise = (const double*)dn;
    }
    if (gate_parallel) {
        void* dq;
        CKL(dev_buf(h, B_QUAT, sizeof(double) * 4 * (size_t)S * (size_t)gate_count, &dq));
        CKL(launch_mc_noise_gate_parallel(h, a, (double*)dq, &nblocks));
    } else {
        CKL(launch_mc_noise(h, a, algo, &nblocks));
    }
    if (d_stats) CKL(launch_stats_finalize(h, a.partials, nblocks, d_stats));
    return RBQ_OK;
}
int rbq_mc_noise_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt, int gate, int64_t gate_count, int64_t S,
                     double fq, const double* d_noise, int64_t noise_len, double noise_dt_inv, const double* d_psi0, int algo,
                     double* d_fid, rbq_stats* d_stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_noise_dev", d_controls, nknots, dt, gate_count, S, algo));
    if (S > 0 && (!d_noise || !d_psi0)) return fail(h, RBQ_E_NULL, "rbq_mc_noise_dev: noise/psi0 is NULL");
    if (noise_len < 1) return fail(h, RBQ_E_SIZE, "rbq_mc_noise_dev: noise_len=%lld", (long long)noise_len);
    return mc_noise_core(h, d_controls, nknots, dt, gate, gate_count, 0, S, fq, d_noise, noise_len, noise_dt_inv, d_psi0, 0, 0,
                         0.0, algo, d_fid, d_stats);
}
int rbq_mc_pink_philox_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                           int64_t first, int64_t S, uint64_t seed, double fq, double namp, double noise_dt_inv, int algo,
                           double* d_fid, rbq_stats* d_stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_pink_philox_dev", d_controls, nknots, dt, gate_count, S, algo));
    if (first < 0) return fail(h, RBQ_E_SIZE, "rbq_mc_pink_philox_dev: first=%lld", (long long)first);
    // same length run_sim_prop gives gen_pink_noise: Int(ceil(evolution_time * noise_dt_inv)) + 1  (spin.jl:1569-1570)
    const int64_t noise_len = (int64_t)ceil((double)nknots * dt * (double)gate_count * noise_dt_inv) + 1;
    return mc_noise_core(h, d_controls, nknots, dt, gate, gate_count, first, S, fq, nullptr, noise_len, noise_dt_inv, nullptr, 1,
                         seed, namp, algo, d_fid, d_stats);
}
int rbq_mc_noise(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count, int64_t S,
                 double fq, const double* noise, int64_t noise_len, double noise_dt_inv, const double* psi0, int algo,
                 double* fid, rbq_stats* stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_noise", controls, nknots, dt, gate_count, S, algo));
    if (S > 0 && (!noise || !psi0)) return fail(h, RBQ_E_NULL, "rbq_mc_noise: noise/psi0 is NULL");
    if (noise_len < 1) return fail(h, RBQ_E_SIZE, "rbq_mc_noise: noise_len=%lld", (long long)noise_len);
    void *dc, *dn, *dpsi, *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc));
    CKL(dev_buf(h, B_IN4, sizeof(double) * S * noise_len, &dn));
    CKL(dev_buf(h, B_IN3, sizeof(double) * S * 4, &dpsi));
    if (fid) CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, &dfid));
    if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst)); }
    CK(cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
    if (S > 0) {
        CK(cudaMemcpyAsync(dn, noise, sizeof(double) * S * noise_len, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dpsi, psi0, sizeof(double) * S * 4, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = mc_noise_core(h, (double*)dc, nknots, dt, gate, gate_count, 0, S, fq, (double*)dn, noise_len, noise_dt_inv,
                           (double*)dpsi, 0, 0, 0.0, algo, (double*)dfid, (rbq_stats*)dst);
    if (rc) return rc;
    if (fid && S > 0) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_mc_pink_philox(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count,
                       int64_t first, int64_t S, uint64_t seed, double fq, double namp, double noise_dt_inv, int algo,
                       double* fid, rbq_stats* stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_mc_pink_philox", controls, nknots, dt, gate_count, S, algo));
    void *dc, *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc));
    if (fid) CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, &dfid));
    if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst)); }
    CK(cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
    int rc = rbq_mc_pink_philox_dev(h, (double*)dc, nknots, dt, gate, gate_count, first, S, seed, fq, namp, noise_dt_inv, algo,
                                    (double*)dfid, (rbq_stats*)dst);
    if (rc) return rc;
    if (fid && S > 0) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_pink_noise(rbq_handle* h, int64_t first, int64_t S, uint64_t seed, double namp, double noise_dt_inv, int64_t noise_len,
                   double* noise) {
    ENTER(h);
    if (!noise) return fail(h, RBQ_E_NULL, "rbq_pink_noise: null pointer");
    if (S < 0 || first < 0 || noise_len < 1) return fail(h, RBQ_E_SIZE, "rbq_pink_noise: bad sizes");
    if (S == 0) return RBQ_OK;
    void* dn;
    CKL(dev_buf(h, B_IN4, sizeof(double) * S * noise_len, &dn));
    CKL(launch_pink_noise(h, first, S, seed, namp, noise_dt_inv, noise_len, (double*)dn));
    CK(cudaMemcpyAsync(noise, dn, sizeof(double) * S * noise_len, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- Lindblad ---------------------------------------------------------------------------------------------------------
int rbq_lindblad_dev(rbq_handle* h, const double* d_controls, int64_t nknots, double dt, int gate, int64_t gate_count, int64_t S,
                     double fq, const double* d_da, const double* d_rho0, int channels, double t2_gamma_c, double t2_gamma_f,
                     double* d_fid, double* d_rho_out, rbq_stats* d_stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_lindblad_dev", d_controls, nknots, dt, gate_count, S, RBQ_ALGO_SU2));
    if (S > 0 && !d_rho0) return fail(h, RBQ_E_NULL, "rbq_lindblad_dev: rho0 is NULL");
    if (channels & ~(RBQ_LINDBLAD_T1 | RBQ_LINDBLAD_T2)) return fail(h, RBQ_E_MODEL, "rbq_lindblad_dev: unknown channel bits %d", channels);
    LindbladArgs a;
    memset(&a, 0, sizeof(a));
    if (!make_gate_tables(gate, nullptr, &a.gate)) return fail(h, RBQ_E_GATE, "rbq_lindblad_dev: unknown gate %d", gate);
    if (S == 0) return RBQ_OK;
    CKL(stage_pulse(h, d_controls, nknots, &a.controls, &a.amax));
    a.nk = nknots; a.dt = dt; a.gate_count = gate_count; a.S = S; a.fq = fq; a.da = d_da; a.rho0 = d_rho0;
    a.t2_gamma_c = t2_gamma_c; a.t2_gamma_f = t2_gamma_f; a.fid = d_fid; a.rho_out = d_rho_out;
    if (d_stats) {
        void* pp;
        CKL(dev_buf(h, B_PARTIALS, sizeof(BlockPartial) * lindblad_blocks(h, S), &pp));
        a.partials = (BlockPartial*)pp;
        a.hist = (unsigned long long*)d_stats->hist;
    }
    int nblocks = 0;
    // no per-realisation amplitude offsets and time-independent rates: the one-gate map is the same for every realisation
    // (the reference recomputes it per seed, figures.jl:194-205); compose it once when it will be reused enough to matter
    const bool time_dep = (channels & RBQ_LINDBLAD_T2) && t2_gamma_f != 0.0;
    if (!d_da && !time_dep && (gate_count > 3 || S >= 64) && h->opt.lindblad_shared_map) {
        void* sc;
        CKL(dev_buf(h, B_SCALARS, 256, &sc));
        CKL(launch_lindblad_shared_map(h, a, channels, (double*)sc + 8));
        a.shared_map = (const double*)sc + 8;
    }
    // per-realisation offsets with time-independent rates: per-knot Frechet table instead of the series (rbq_mc.cu)
    if (d_da && !time_dep && h->opt.lindblad_table && nknots <= ((int64_t)1 << 24)) {
        void* tb;
        CKL(dev_buf(h, B_QUAT, sizeof(double) * 48 * (size_t)nknots, &tb));
        CKL(launch_lindblad_table(h, a, channels, d_controls, (double*)tb, &nblocks));
    } else {
        CKL(launch_lindblad(h, a, channels, &nblocks));
    }
    if (d_stats) CKL(launch_stats_finalize(h, a.partials, nblocks, d_stats));
    return RBQ_OK;
}
int rbq_lindblad(rbq_handle* h, const double* controls, int64_t nknots, double dt, int gate, int64_t gate_count, int64_t S,
                 double fq, const double* da, const double* rho0, int channels, double t2_gamma_c, double t2_gamma_f, double* fid,
                 double* rho_out, rbq_stats* stats) {
    ENTER(h);
    CKL(check_mc(h, "rbq_lindblad", controls, nknots, dt, gate_count, S, RBQ_ALGO_SU2));
    if (S > 0 && !rho0) return fail(h, RBQ_E_NULL, "rbq_lindblad: rho0 is NULL");
    void *dc, *dda = nullptr, *drho, *dfid = nullptr, *dro = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nknots, &dc));
    CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda));
    CKL(dev_buf(h, B_IN3, sizeof(double) * S * 8, &drho));
    if (fid) CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, &dfid));
    if (rho_out) CKL(dev_buf(h, B_OUT1, sizeof(double) * S * 8, &dro));
    if (stats) { CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), &dst)); CKL(rbq_stats_reset_dev(h, (rbq_stats*)dst)); }
    CK(cudaMemcpyAsync(dc, controls, sizeof(double) * nknots, cudaMemcpyHostToDevice, h->stream));
    if (S > 0) {
        if (da) CK(cudaMemcpyAsync(dda, da, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(drho, rho0, sizeof(double) * S * 8, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = rbq_lindblad_dev(h, (double*)dc, nknots, dt, gate, gate_count, S, fq, da ? (double*)dda : nullptr, (double*)drho,
                              channels, t2_gamma_c, t2_gamma_f, (double*)dfid, (double*)dro, (rbq_stats*)dst);
    if (rc) return rc;
    if (fid && S > 0) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    if (rho_out && S > 0) CK(cudaMemcpyAsync(rho_out, dro, sizeof(double) * S * 8, cudaMemcpyDeviceToHost, h->stream));
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- analytic gates (per-segment durations) ----------------------------------------------------------------------------
static int segments_common(rbq_handle* h, const char* who, const double* amps, const double* durs, int64_t nseg, int gate,
                           int64_t gate_count, int64_t S, SegmentArgs* a, void** d_fid, void** d_st, double* fid, rbq_stats* stats,
                           int64_t G1) {
    if (!amps || !durs) return fail(h, RBQ_E_NULL, "%s: amps/durs is NULL", who);
    if (nseg < 1 || nseg > (1 << 24) || gate_count < 0 || S < 0) return fail(h, RBQ_E_SIZE, "%s: nseg=%lld gate_count=%lld S=%lld", who, (long long)nseg, (long long)gate_count, (long long)S);
    memset(a, 0, sizeof(*a));
    if (!make_gate_tables(gate, &a->gate, &a->rot)) return fail(h, RBQ_E_GATE, "%s: unknown gate %d", who, gate);
    void *da_, *dd_;
    CKL(dev_buf(h, B_IN0, sizeof(double) * nseg, &da_)); CKL(dev_buf(h, B_IN5, sizeof(double) * nseg, &dd_));
    CK(cudaMemcpyAsync(da_, amps, sizeof(double) * nseg, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dd_, durs, sizeof(double) * nseg, cudaMemcpyHostToDevice, h->stream));
    a->amps = (double*)da_; a->durs = (double*)dd_; a->nseg = nseg; a->gate_count = gate_count; a->S = S;
    if (fid) { CKL(dev_buf(h, B_OUT0, sizeof(double) * S * G1, d_fid)); a->fid = (double*)*d_fid; }
    if (stats) {
        CKL(dev_buf(h, B_STATS, sizeof(rbq_stats), d_st));
        CKL(rbq_stats_reset_dev(h, (rbq_stats*)*d_st));
        void* pp;
        CKL(dev_buf(h, B_PARTIALS, sizeof(BlockPartial) * (size_t)((S + 255) / 256 + 1), &pp));
        a->partials = (BlockPartial*)pp;
        a->hist = (unsigned long long*)((rbq_stats*)*d_st)->hist;
    }
    return 0;
}
int rbq_mc_static_segments(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, int gate, int64_t gate_count,
                           int64_t S, const double* fq, const double* da, const double* psi0, double* fid, rbq_stats* stats) {
    ENTER(h);
    SegmentArgs a;
    void *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(segments_common(h, "rbq_mc_static_segments", amps, durs, nseg, gate, gate_count, S, &a, &dfid, &dst, fid, stats, G1));
    if (S > 0) {
        if (!fq || !psi0) return fail(h, RBQ_E_NULL, "rbq_mc_static_segments: fq/psi0 is NULL");
        void *dfq, *dda, *dpsi;
        CKL(dev_buf(h, B_IN1, sizeof(double) * S, &dfq)); CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda));
        CKL(dev_buf(h, B_IN3, sizeof(double) * 4 * S, &dpsi));
        CK(cudaMemcpyAsync(dfq, fq, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        if (da) CK(cudaMemcpyAsync(dda, da, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dpsi, psi0, sizeof(double) * 4 * S, cudaMemcpyHostToDevice, h->stream));
        a.fq = (double*)dfq; a.da = da ? (double*)dda : nullptr; a.psi0 = (double*)dpsi;
        int nblocks = 0;
        CKL(launch_segments(h, a, false, &nblocks));
        if (stats) CKL(launch_stats_finalize(h, a.partials, nblocks, (rbq_stats*)dst));
        if (fid) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_mc_noise_segments(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg, int gate,
                          int64_t gate_count, int64_t S, double fq, const double* noise, int64_t noise_len, const double* psi0,
                          double* fid, rbq_stats* stats) {
    ENTER(h);
    SegmentArgs a;
    void *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(segments_common(h, "rbq_mc_noise_segments", amps, durs, nseg, gate, gate_count, S, &a, &dfid, &dst, fid, stats, G1));
    if (!noise_idx) return fail(h, RBQ_E_NULL, "rbq_mc_noise_segments: noise_idx is NULL");
    if (noise_len < 1) return fail(h, RBQ_E_SIZE, "rbq_mc_noise_segments: noise_len=%lld", (long long)noise_len);
    for (int64_t j = 0; j < nseg; ++j)
        if (noise_idx[j] < 0 || noise_idx[j] >= noise_len)
            return fail(h, RBQ_E_SIZE, "rbq_mc_noise_segments: noise_idx[%lld]=%d outside [0,%lld)", (long long)j, noise_idx[j], (long long)noise_len);
    if (S > 0) {
        if (!noise || !psi0) return fail(h, RBQ_E_NULL, "rbq_mc_noise_segments: noise/psi0 is NULL");
        void *didx, *dn, *dpsi;
        CKL(dev_buf(h, B_IN1, sizeof(int32_t) * nseg, &didx)); CKL(dev_buf(h, B_IN4, sizeof(double) * S * noise_len, &dn));
        CKL(dev_buf(h, B_IN3, sizeof(double) * 4 * S, &dpsi));
        CK(cudaMemcpyAsync(didx, noise_idx, sizeof(int32_t) * nseg, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dn, noise, sizeof(double) * S * noise_len, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dpsi, psi0, sizeof(double) * 4 * S, cudaMemcpyHostToDevice, h->stream));
        a.fq = nullptr; a.fq_scalar = fq; a.noise = (double*)dn; a.nidx = (int32_t*)didx; a.noise_len = noise_len; a.psi0 = (double*)dpsi;
        int nblocks = 0;
        CKL(launch_segments(h, a, false, &nblocks));
        if (stats) CKL(launch_stats_finalize(h, a.partials, nblocks, (rbq_stats*)dst));
        if (fid) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_mc_timeline(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg,
                    const int64_t* save_after, int64_t nsave, int gate, int64_t S, const double* fq, const double* da,
                    const double* noise, int64_t noise_len, const double* psi0, double* fid, rbq_stats* stats) {
    return rbq_mc_timeline_states(h, amps, durs, noise_idx, nseg, save_after, nsave, gate, S, fq, da, noise, noise_len, psi0, fid, nullptr, stats);
}
int rbq_mc_timeline_states(rbq_handle* h, const double* amps, const double* durs, const int32_t* noise_idx, int64_t nseg,
                           const int64_t* save_after, int64_t nsave, int gate, int64_t S, const double* fq, const double* da,
                           const double* noise, int64_t noise_len, const double* psi0, double* fid, double* states, rbq_stats* stats) {
    ENTER(h);
    SegmentArgs a;
    void *dfid = nullptr, *dst = nullptr;
    if (nsave < 0 || !save_after) return fail(h, nsave < 0 ? RBQ_E_SIZE : RBQ_E_NULL, "rbq_mc_timeline: save_after/nsave");
    CKL(segments_common(h, "rbq_mc_timeline", amps, durs, nseg, gate, nsave, S, &a, &dfid, &dst, fid, stats, nsave + 1));
    int64_t prev = 0;
    for (int64_t g = 0; g < nsave; ++g) {
        if (save_after[g] < prev || save_after[g] > nseg) return fail(h, RBQ_E_SIZE, "rbq_mc_timeline: save_after[%lld]=%lld not ascending within [0,%lld]", (long long)g, (long long)save_after[g], (long long)nseg);
        prev = save_after[g];
    }
    if ((noise != nullptr) != (noise_idx != nullptr)) return fail(h, RBQ_E_NULL, "rbq_mc_timeline: noise and noise_idx go together");
    if (noise) {
        if (noise_len < 1) return fail(h, RBQ_E_SIZE, "rbq_mc_timeline: noise_len=%lld", (long long)noise_len);
        for (int64_t j = 0; j < nseg; ++j)
            if (noise_idx[j] < 0 || noise_idx[j] >= noise_len)
                return fail(h, RBQ_E_SIZE, "rbq_mc_timeline: noise_idx[%lld]=%d outside [0,%lld)", (long long)j, noise_idx[j], (long long)noise_len);
    }
    if (S > 0) {
        if (!fq || !psi0) return fail(h, RBQ_E_NULL, "rbq_mc_timeline: fq/psi0 is NULL");
        void *dfq, *dda, *dpsi, *didx = nullptr, *dn = nullptr, *dsave;
        CKL(dev_buf(h, B_IN1, sizeof(double) * S, &dfq)); CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda));
        CKL(dev_buf(h, B_IN3, sizeof(double) * 4 * S, &dpsi));
        CKL(dev_buf(h, B_QUAT, sizeof(int64_t) * (size_t)(nsave > 0 ? nsave : 1), &dsave));
        CK(cudaMemcpyAsync(dfq, fq, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        if (da) CK(cudaMemcpyAsync(dda, da, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dpsi, psi0, sizeof(double) * 4 * S, cudaMemcpyHostToDevice, h->stream));
        if (nsave > 0) CK(cudaMemcpyAsync(dsave, save_after, sizeof(int64_t) * nsave, cudaMemcpyHostToDevice, h->stream));
        if (noise) {
            CKL(dev_buf(h, B_NOISE, sizeof(int32_t) * nseg, &didx)); CKL(dev_buf(h, B_IN4, sizeof(double) * S * noise_len, &dn));
            CK(cudaMemcpyAsync(didx, noise_idx, sizeof(int32_t) * nseg, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(dn, noise, sizeof(double) * S * noise_len, cudaMemcpyHostToDevice, h->stream));
        }
        a.fq = (double*)dfq; a.da = da ? (double*)dda : nullptr; a.psi0 = (double*)dpsi;
        a.noise = (double*)dn; a.nidx = (int32_t*)didx; a.noise_len = noise_len;
        void* dstates = nullptr;
        if (states) { CKL(dev_buf(h, B_OUT1, sizeof(double) * 4 * S * (nsave + 1), &dstates)); a.states_out = (double*)dstates; }
        int nblocks = 0;
        CKL(launch_timeline(h, a, (const int64_t*)dsave, nsave, &nblocks));
        if (stats) CKL(launch_stats_finalize(h, a.partials, nblocks, (rbq_stats*)dst));
        if (fid) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * (nsave + 1), cudaMemcpyDeviceToHost, h->stream));
        if (states) CK(cudaMemcpyAsync(states, dstates, sizeof(double) * 4 * S * (nsave + 1), cudaMemcpyDeviceToHost, h->stream));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_lindblad_timeline(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, const int64_t* save_after,
                          int64_t nsave, int gate, int64_t S, double fq, const double* da, const double* rho0, int channels,
                          double t2_gamma_c, double t2_gamma_f, double* fid, double* rho_saves, rbq_stats* stats) {
    ENTER(h);
    if (channels & ~(RBQ_LINDBLAD_T1 | RBQ_LINDBLAD_T2)) return fail(h, RBQ_E_MODEL, "rbq_lindblad_timeline: unknown channel bits %d", channels);
    SegmentArgs a;
    void *dfid = nullptr, *dst = nullptr;
    if (nsave < 0 || !save_after) return fail(h, nsave < 0 ? RBQ_E_SIZE : RBQ_E_NULL, "rbq_lindblad_timeline: save_after/nsave");
    CKL(segments_common(h, "rbq_lindblad_timeline", amps, durs, nseg, gate, nsave, S, &a, &dfid, &dst, fid, stats, nsave + 1));
    int64_t prev = 0;
    for (int64_t g = 0; g < nsave; ++g) {
        if (save_after[g] < prev || save_after[g] > nseg) return fail(h, RBQ_E_SIZE, "rbq_lindblad_timeline: save_after[%lld]=%lld not ascending within [0,%lld]", (long long)g, (long long)save_after[g], (long long)nseg);
        prev = save_after[g];
    }
    if (S > 0) {
        if (!rho0) return fail(h, RBQ_E_NULL, "rbq_lindblad_timeline: rho0 is NULL");
        void *dda, *drho, *dsave, *dsaves = nullptr;
        CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda)); CKL(dev_buf(h, B_IN3, sizeof(double) * 8 * S, &drho));
        CKL(dev_buf(h, B_QUAT, sizeof(int64_t) * (size_t)(nsave > 0 ? nsave : 1), &dsave));
        if (da) CK(cudaMemcpyAsync(dda, da, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(drho, rho0, sizeof(double) * 8 * S, cudaMemcpyHostToDevice, h->stream));
        if (nsave > 0) CK(cudaMemcpyAsync(dsave, save_after, sizeof(int64_t) * nsave, cudaMemcpyHostToDevice, h->stream));
        if (rho_saves) CKL(dev_buf(h, B_OUT1, sizeof(double) * 8 * S * (nsave + 1), &dsaves));
        a.fq_scalar = fq; a.da = da ? (double*)dda : nullptr; a.rho0 = (double*)drho; a.rho_saves = (double*)dsaves;
        a.channels = channels; a.t2_gamma_c = (channels & RBQ_LINDBLAD_T2) ? t2_gamma_c : 0.0;
        a.t2_gamma_f = (channels & RBQ_LINDBLAD_T2) ? t2_gamma_f : 0.0;
        int nblocks = 0;
        CKL(launch_timeline(h, a, (const int64_t*)dsave, nsave, &nblocks));
        if (stats) CKL(launch_stats_finalize(h, a.partials, nblocks, (rbq_stats*)dst));
        if (fid) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * (nsave + 1), cudaMemcpyDeviceToHost, h->stream));
        if (rho_saves) CK(cudaMemcpyAsync(rho_saves, dsaves, sizeof(double) * 8 * S * (nsave + 1), cudaMemcpyDeviceToHost, h->stream));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}
int rbq_lindblad_segments(rbq_handle* h, const double* amps, const double* durs, int64_t nseg, int gate, int64_t gate_count,
                          int64_t S, double fq, const double* da, const double* rho0, int channels, double t2_gamma_c, double* fid,
                          double* rho_out, rbq_stats* stats) {
    ENTER(h);
    if (channels & ~(RBQ_LINDBLAD_T1 | RBQ_LINDBLAD_T2)) return fail(h, RBQ_E_MODEL, "rbq_lindblad_segments: unknown channel bits %d", channels);
    SegmentArgs a;
    void *dfid = nullptr, *dst = nullptr;
    const int64_t G1 = gate_count + 1;
    CKL(segments_common(h, "rbq_lindblad_segments", amps, durs, nseg, gate, gate_count, S, &a, &dfid, &dst, fid, stats, G1));
    if (S > 0) {
        if (!rho0) return fail(h, RBQ_E_NULL, "rbq_lindblad_segments: rho0 is NULL");
        void *dda, *drho, *dro = nullptr;
        CKL(dev_buf(h, B_IN2, sizeof(double) * S, &dda)); CKL(dev_buf(h, B_IN3, sizeof(double) * 8 * S, &drho));
        if (da) CK(cudaMemcpyAsync(dda, da, sizeof(double) * S, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(drho, rho0, sizeof(double) * 8 * S, cudaMemcpyHostToDevice, h->stream));
        if (rho_out) CKL(dev_buf(h, B_OUT1, sizeof(double) * 8 * S, &dro));
        a.fq_scalar = fq; a.da = da ? (double*)dda : nullptr; a.rho0 = (double*)drho; a.rho_out = (double*)dro;
        a.channels = channels; a.t2_gamma_c = (channels & RBQ_LINDBLAD_T2) ? t2_gamma_c : 0.0;
        int nblocks = 0;
        CKL(launch_segments(h, a, true, &nblocks));
        if (stats) CKL(launch_stats_finalize(h, a.partials, nblocks, (rbq_stats*)dst));
        if (fid) CK(cudaMemcpyAsync(fid, dfid, sizeof(double) * S * G1, cudaMemcpyDeviceToHost, h->stream));
        if (rho_out) CK(cudaMemcpyAsync(rho_out, dro, sizeof(double) * 8 * S, cudaMemcpyDeviceToHost, h->stream));
    }
    if (stats) CK(cudaMemcpyAsync(stats, dst, sizeof(rbq_stats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- iLQR backward pass ------------------------------------------------------------------------------------------------
static int check_bp(rbq_handle* h, const char* who, int n, int m, int64_t N, int64_t B, int reg_type) {
    if (n < 1 || n > 64 || m < 1 || m > 4 || N < 1 || B < 0) return fail(h, RBQ_E_SIZE, "%s: n=%d m=%d N=%lld B=%lld (n <= 64, m <= 4)", who, n, m, (long long)N, (long long)B);
    if (reg_type != RBQ_REG_CONTROL && reg_type != RBQ_REG_STATE) return fail(h, RBQ_E_MODEL, "%s: unknown reg_type %d", who, reg_type);
    return 0;
}
int rbq_backward_pass_dev(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_AB, const double* d_Exx,
                          const double* d_Ex, const double* d_Euu, const double* d_Eux, const double* d_Eu, double rho,
                          int reg_type, double* d_K, double* d_d, double* d_dV, int64_t* d_fail) {
    ENTER(h);
    CKL(check_bp(h, "rbq_backward_pass_dev", n, m, N, B, reg_type));
    if (B == 0) return RBQ_OK;
    if (!d_Exx || !d_Ex || !d_dV || !d_fail || (N > 1 && (!d_AB || !d_Euu || !d_Eux || !d_Eu || !d_K || !d_d)))
        return fail(h, RBQ_E_NULL, "rbq_backward_pass_dev: null pointer");
    CKL(launch_backward_pass(h, n, m, N, B, d_AB, d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, rho, reg_type, d_K, d_d, d_dV, (long long*)d_fail));
    return RBQ_OK;
}
int rbq_backward_pass(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* AB, const double* Exx, const double* Ex,
                      const double* Euu, const double* Eux, const double* Eu, double rho, int reg_type, double* K, double* d,
                      double* dV, int64_t* failed) {
    ENTER(h);
    CKL(check_bp(h, "rbq_backward_pass", n, m, N, B, reg_type));
    if (B == 0) return RBQ_OK;
    if (!Exx || !Ex || !dV || !failed || (N > 1 && (!AB || !Euu || !Eux || !Eu || !K || !d)))
        return fail(h, RBQ_E_NULL, "rbq_backward_pass: null pointer");
    const size_t nm = (size_t)n + m, K1 = (size_t)(N - 1) * B, KN = (size_t)N * B;
    const size_t bAB = sizeof(double) * K1 * n * nm, bExx = sizeof(double) * KN * n * n, bEx = sizeof(double) * KN * n;
    const size_t bEuu = sizeof(double) * K1 * m * m, bEux = sizeof(double) * K1 * m * n, bEu = sizeof(double) * K1 * m;
    void *dAB, *dExx, *dEx, *dEuu, *dEux, *dEu, *dK, *dd, *dsm;
    CKL(dev_buf(h, B_IN0, bAB, &dAB)); CKL(dev_buf(h, B_IN1, bExx, &dExx)); CKL(dev_buf(h, B_IN2, bEx, &dEx));
    CKL(dev_buf(h, B_IN3, bEuu, &dEuu)); CKL(dev_buf(h, B_IN4, bEux, &dEux)); CKL(dev_buf(h, B_IN5, bEu, &dEu));
    CKL(dev_buf(h, B_OUT0, bEux, &dK)); CKL(dev_buf(h, B_OUT1, bEu, &dd));
    CKL(dev_buf(h, B_SCALARS, 256 + 24 * (size_t)B, &dsm));
    double* ddV = (double*)dsm + 32; long long* dfail = (long long*)((double*)dsm + 32 + 2 * B);
    if (N > 1) {
        CK(cudaMemcpyAsync(dAB, AB, bAB, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dEuu, Euu, bEuu, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dEux, Eux, bEux, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dEu, Eu, bEu, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(dExx, Exx, bExx, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dEx, Ex, bEx, cudaMemcpyHostToDevice, h->stream));
    CKL(launch_backward_pass(h, n, m, N, B, (double*)dAB, (double*)dExx, (double*)dEx, (double*)dEuu, (double*)dEux, (double*)dEu, rho,
                             reg_type, (double*)dK, (double*)dd, ddV, dfail));
    if (N > 1) {
        CK(cudaMemcpyAsync(K, dK, bEux, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(d, dd, bEu, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaMemcpyAsync(dV, ddV, sizeof(double) * 2 * B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(failed, dfail, sizeof(int64_t) * B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- augmented-Lagrangian cost expansion (device resident) ---------------------------------------------------------------
int rbq_al_cost_expansion_dev(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* d_X, const double* d_U,
                              const double* d_Qdiag, const double* d_Rdiag, const double* d_Qfdiag, const double* d_xf,
                              const double* d_w, int ncon, const rbq_constraint* d_cons, const int32_t* d_idx_pool,
                              const double* d_val_pool, const double* d_lambda, const double* d_mu, int64_t dual_stride,
                              double* d_Exx, double* d_Ex, double* d_Euu, double* d_Eux, double* d_Eu, double* d_c, double* d_J) {
    ENTER(h);
    if (n < 1 || m < 1 || N < 1 || B < 0 || B > 65535 || ncon < 0 || dual_stride < 0)
        return fail(h, RBQ_E_SIZE, "rbq_al_cost_expansion_dev: n=%d m=%d N=%lld B=%lld ncon=%d", n, m, (long long)N, (long long)B, ncon);
    if (B == 0) return RBQ_OK;
    if (!d_X || (N > 1 && !d_U) || !d_Qdiag || !d_Rdiag || !d_Qfdiag || !d_xf) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion_dev: null pointer");
    if (ncon > 0 && (!d_cons || !d_val_pool || !d_lambda || !d_mu)) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion_dev: null constraint data");
    if (d_Exx && (!d_Ex || (N > 1 && (!d_Euu || !d_Eux || !d_Eu)))) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion_dev: Exx given without the other blocks");
    CKL(launch_al_cost(h, n, m, N, B, dual_stride, d_X, d_U, d_Qdiag, d_Rdiag, d_Qfdiag, d_xf, d_w, ncon, d_cons, d_idx_pool, d_val_pool, d_lambda,
                       d_mu, d_Exx, d_Ex, d_Euu, d_Eux, d_Eu, d_c, d_J));
    return RBQ_OK;
}

// outer-loop updates on the device: multipliers, penalties, largest violation
int rbq_al_dual_update_dev(rbq_handle* h, int64_t N, int64_t B, int ncon, const rbq_constraint* d_cons, int64_t max_p,
                           const double* d_c, double* d_lambda, double* d_mu, int64_t dual_stride, double phi, double mu_max,
                           double lambda_max, int update, double* d_max_violation) {
    ENTER(h);
    if (N < 1 || B < 0 || ncon < 0 || max_p < 1 || dual_stride < 0) return fail(h, RBQ_E_SIZE, "rbq_al_dual_update_dev: N=%lld B=%lld ncon=%d max_p=%lld", (long long)N, (long long)B, ncon, (long long)max_p);
    if (!d_max_violation) return fail(h, RBQ_E_NULL, "rbq_al_dual_update_dev: null pointer");
    if (ncon == 0 || B == 0) { CK(cudaMemsetAsync(d_max_violation, 0, sizeof(double), h->stream)); return RBQ_OK; }
    if (!d_cons || !d_c || (update && (!d_lambda || !d_mu))) return fail(h, RBQ_E_NULL, "rbq_al_dual_update_dev: null pointer");
    CKL(launch_al_dual_update(h, ncon, d_cons, N * max_p, B, dual_stride, d_c, d_lambda, d_mu, phi, mu_max, lambda_max, update, d_max_violation));
    return RBQ_OK;
}

// host-pointer form: every input is packed into one device arena, every output into another
int rbq_al_cost_expansion(rbq_handle* h, int n, int m, int64_t N, int64_t B, const double* X, const double* U, const double* Qdiag,
                          const double* Rdiag, const double* Qfdiag, const double* xf, const double* w, int ncon,
                          const rbq_constraint* cons, const int32_t* idx_pool, int64_t nidx, const double* val_pool, int64_t nval,
                          const double* lambda, const double* mu, int64_t dual_stride, double* Exx, double* Ex, double* Euu,
                          double* Eux, double* Eu, double* c, double* J) {
    ENTER(h);
    if (n < 1 || m < 1 || N < 1 || B < 0 || B > 65535 || ncon < 0 || dual_stride < 0 || nidx < 0 || nval < 0)
        return fail(h, RBQ_E_SIZE, "rbq_al_cost_expansion: n=%d m=%d N=%lld B=%lld ncon=%d", n, m, (long long)N, (long long)B, ncon);
    if (B == 0) return RBQ_OK;
    if (!X || (N > 1 && !U) || !Qdiag || !Rdiag || !Qfdiag || !xf) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion: null pointer");
    if (ncon > 0 && (!cons || !val_pool || !lambda || !mu || (nidx > 0 && !idx_pool))) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion: null constraint data");
    if (Exx && (!Ex || (N > 1 && (!Euu || !Eux || !Eu)))) return fail(h, RBQ_E_NULL, "rbq_al_cost_expansion: Exx given without the other blocks");
    int64_t dual_size = 0;
    for (int i = 0; i < ncon; ++i) {
        const int64_t end = cons[i].dual_off + (cons[i].k1 - cons[i].k0) * cons[i].p;
        if (cons[i].dual_off < dual_size || cons[i].k0 < 0 || cons[i].k1 > N || cons[i].k1 < cons[i].k0)
            return fail(h, RBQ_E_SIZE, "rbq_al_cost_expansion: constraint %d: rows must be in dual_off order with knots inside [0, N]", i);
        dual_size = end;
    }
    if (dual_stride != 0 && dual_stride < dual_size) return fail(h, RBQ_E_SIZE, "rbq_al_cost_expansion: dual_stride too small");
    const int64_t nd = dual_stride ? dual_stride * (B - 1) + dual_size : dual_size;     // doubles of lambda / mu read
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // arena layout (bytes)
    const size_t sX = sizeof(double) * B * N * n, sU = sizeof(double) * B * (N - 1) * m, sn = sizeof(double) * n, sm_ = sizeof(double) * m;
    const size_t sW = w ? sizeof(double) * N : 0, sC = sizeof(rbq_constraint) * ncon, sI = sizeof(int32_t) * nidx, sV = sizeof(double) * nval;
    const size_t sD = sizeof(double) * nd;
    size_t off = 0;
    auto take = [&](size_t b) { const size_t o = off; off += up(b ? b : 16); return o; };
    const size_t oX = take(sX), oU = take(sU), oQ = take(sn), oR = take(sm_), oQf = take(sn), oxf = take(sn), oW = take(sW), oC = take(sC),
                 oI = take(sI), oV = take(sV), oL = take(sD), oM = take(sD);
    void* in; CKL(dev_buf(h, B_IN0, off, &in));
    char* ib = (char*)in;
    auto put = [&](size_t o, const void* src, size_t b) -> cudaError_t { return b && src ? cudaMemcpyAsync(ib + o, src, b, cudaMemcpyHostToDevice, h->stream) : cudaSuccess; };
    CK(put(oX, X, sX)); CK(put(oU, U, sU)); CK(put(oQ, Qdiag, sn)); CK(put(oR, Rdiag, sm_)); CK(put(oQf, Qfdiag, sn)); CK(put(oxf, xf, sn));
    CK(put(oW, w, sW)); CK(put(oC, cons, sC)); CK(put(oI, idx_pool, sI)); CK(put(oV, val_pool, sV)); CK(put(oL, lambda, sD)); CK(put(oM, mu, sD));
    const size_t sExx = Exx ? sizeof(double) * B * N * n * n : 0, sEx = Exx ? sizeof(double) * B * N * n : 0;
    const size_t sEuu = Exx ? sizeof(double) * B * (N - 1) * m * m : 0, sEux = Exx ? sizeof(double) * B * (N - 1) * m * n : 0;
    const size_t sEu = Exx ? sizeof(double) * B * (N - 1) * m : 0, sc = c ? sizeof(double) * B * dual_size : 0, sJ = J ? sizeof(double) * B * N : 0;
    off = 0;
    const size_t oExx = take(sExx), oEx = take(sEx), oEuu = take(sEuu), oEux = take(sEux), oEu = take(sEu), oc = take(sc), oJ = take(sJ);
    void* out; CKL(dev_buf(h, B_OUT0, off, &out));
    char* ob = (char*)out;
    auto dp = [&](size_t o, const void* host) -> double* { return host ? (double*)(ob + o) : nullptr; };
    CKL(launch_al_cost(h, n, m, N, B, dual_stride, (double*)(ib + oX), (double*)(ib + oU), (double*)(ib + oQ), (double*)(ib + oR),
                       (double*)(ib + oQf), (double*)(ib + oxf), w ? (double*)(ib + oW) : nullptr, ncon, (rbq_constraint*)(ib + oC),
                       (int32_t*)(ib + oI), (double*)(ib + oV), (double*)(ib + oL), (double*)(ib + oM), dp(oExx, Exx), dp(oEx, Exx),
                       dp(oEuu, Exx), dp(oEux, Exx), dp(oEu, Exx), dp(oc, c), dp(oJ, J)));
    auto get = [&](void* dst, size_t o, size_t b) -> cudaError_t { return b && dst ? cudaMemcpyAsync(dst, ob + o, b, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess; };
    CK(get(Exx, oExx, sExx)); CK(get(Ex, oEx, sEx)); CK(get(Euu, oEuu, sEuu)); CK(get(Eux, oEux, sEux)); CK(get(Eu, oEu, sEu));
    CK(get(c, oc, sc)); CK(get(J, oJ, sJ));
    CK(cudaStreamSynchronize(h->stream));
    return RBQ_OK;
}

// ---- FP64 roofline denominator -------------------------------------------------------------------------------------------
int rbq_fp64_peak(rbq_handle* h, int iters, double* tflops) {
    ENTER(h);
    if (!tflops) return fail(h, RBQ_E_NULL, "rbq_fp64_peak: null pointer");
    if (iters < 1) return fail(h, RBQ_E_SIZE, "rbq_fp64_peak: iters=%d", iters);
    void* sink;
    CKL(dev_buf(h, B_SCALARS, 256, &sink));
    int nthreads = 0, fpi = 0;
    CKL(launch_fp64_peak(h, iters / 8 + 1, (double*)sink + 4, &nthreads, &fpi));   // warm-up
    CK(cudaEventRecord(h->ev[6], h->stream));
    CKL(launch_fp64_peak(h, iters, (double*)sink + 4, &nthreads, &fpi));
    CK(cudaEventRecord(h->ev[7], h->stream));
    CK(cudaEventSynchronize(h->ev[7]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]));
    *tflops = (double)nthreads * (double)iters * (double)fpi / ((double)ms * 1e-3) * 1e-12;
    return RBQ_OK;
}

}  // extern "C"
