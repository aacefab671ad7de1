// C ABI of libspecfab_b200.so (declared in include/specfab_b200.h).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "sfb_common.cuh"
#include "specfab_b200.h"

struct SfbStepEntry {
    int L, ddrx, R, TN, dfma_node;
    sfb_step_launch_fn fn;
};
#include "gen/registry.inc"

namespace {

struct State {
    int L = 0, n = 0;
    SfbRegConst reg;
    std::string err;
    std::string info;
    std::mutex mu;
} g;

thread_local std::string t_err;

int fail(int code, const std::string& msg) {
    t_err = msg;
    g.err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(SFB_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// src/include/regcalib.f90:1-36 (expo, nu) for even L = 4..20
const double kExpo[9] = {1.700, 1.150, 1.600, 2.000, 2.000, 2.000, 2.500, 2.500, 3.000};
const double kNu[9] = {1.9879322126397958, 3.0011508426238862, 5.7498069921352384, 10.7048905312159288,
                       10.6068117205577668, 13.3591023418822363, 15.3094482670021108, 16.4844589176829217,
                       19.9467342880730136};

const SfbStepEntry* find_step(int L, int ddrx) {
    for (const auto& e : kStepRegistry)
        if (e.L == L && e.ddrx == ddrx) return &e;
    return nullptr;
}

// ---- host-pointer staging: per-thread ring of device chunk buffers -------------------------
struct Slot {
    cudaStream_t st = nullptr;
    double *nin = nullptr, *nout = nullptr, *ug = nullptr, *tau = nullptr, *g0 = nullptr, *lam = nullptr;
    size_t cap_n = 0;   // bytes allocated for nin/nout
    size_t cap_c = 0;   // nodes allocated for ug/tau/g0/lam
};
struct Staging {
    static const int kSlots = 3;
    Slot s[kSlots];
    int dev = -1;
    void release() {
        for (auto& x : s) {
            if (x.st) cudaStreamDestroy(x.st);
            cudaFree(x.nin); cudaFree(x.nout); cudaFree(x.ug); cudaFree(x.tau); cudaFree(x.g0); cudaFree(x.lam);
            x = Slot();
        }
        dev = -1;
    }
};
Staging g_stage;

}  // namespace

extern "C" {

const char* sfb_last_error(void) { return t_err.empty() ? g.err.c_str() : t_err.c_str(); }

int sfb_init(int L) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (L % 2 != 0 || L < 4 || L > SFB_MAXL) return fail(SFB_EINVAL, "sfb_init: L must be even, 4 <= L <= 20");
    if (!find_step(L, 0) || !find_step(L, 1)) return fail(SFB_ENOTBUILT, "sfb_init: kernels for this L were not compiled");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SFB_ECUDA, std::string("sfb_init: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    g.L = L;
    g.n = (L + 1) * (L + 2) / 2;
    const int q = L / 2 - 2;
    g.reg.nu = kNu[q];
    for (int t = 0; t < SFB_NLREG; ++t) g.reg.regdiag[t] = 0.0;
    for (int l = 0; l <= L; l += 2) {
        // abs( Ldiag/(L(L+1)) )**expo   src/dynamics.f90:512  (host libm pow, as the reference)
        const double ldiag = -(double)(l * (l + 1));
        g.reg.regdiag[l / 2] = pow(fabs(ldiag / (double)(L * (L + 1))), kExpo[q]);
    }
    return SFB_OK;
}

int sfb_nlm_len(void) { return g.n; }

int sfb_get_lm(int32_t* lm) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (!lm) return fail(SFB_EINVAL, "null lm");
    int j = 0;
    for (int l = 0; l <= g.L; l += 2)
        for (int m = -l; m <= l; ++m) { lm[2 * j] = l; lm[2 * j + 1] = m; ++j; }
    return SFB_OK;
}

void sfb_finalize(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    g_stage.release();
    g.L = 0; g.n = 0;
}

const char* sfb_build_info(void) {
    if (g.info.empty()) {
        std::string s = "{\"arch\":\"sm_100a\",\"step_kernels\":[";
        bool first = true;
        for (const auto& e : kStepRegistry) {
            char b[160];
            snprintf(b, sizeof b, "%s{\"L\":%d,\"ddrx\":%d,\"roles\":%d,\"tile\":%d,\"dfma_per_node_rhs\":%d}", first ? "" : ",", e.L, e.ddrx, e.R, e.TN, e.dfma_node);
            s += b;
            first = false;
        }
        s += "]}";
        g.info = s;
    }
    return g.info.c_str();
}

static int check_opts(const sfb_step_opts* o) {
    if (!o) return fail(SFB_EINVAL, "null opts");
    if (o->scheme != SFB_EULER && o->scheme != SFB_RK4) return fail(SFB_EINVAL, "scheme must be SFB_EULER or SFB_RK4");
    if (o->nsteps < 1) return fail(SFB_EINVAL, "nsteps must be >= 1");
    if (o->terms & ~(SFB_LROT | SFB_DDRX | SFB_CDRX | SFB_REG)) return fail(SFB_EINVAL, "unknown bits in terms");
    return SFB_OK;
}

int sfb_step_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                     const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                     const sfb_step_opts* o, void* stream) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    int rc = check_opts(o);
    if (rc) return rc;
    if (N < 0 || ld_in < N || ld_out < N || ld_u < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!nlm_in || !nlm_out || !ugrad) return fail(SFB_EINVAL, "null array");
    if (((uintptr_t)nlm_in & 15) || ((uintptr_t)nlm_out & 15)) return fail(SFB_EINVAL, "nlm arrays must be 16-byte aligned");
    const int ddrx = (o->terms & SFB_DDRX) ? 1 : 0;
    if (ddrx && tau && ld_t < N) return fail(SFB_EINVAL, "ld_t < N");
    const SfbStepEntry* ent = find_step(g.L, ddrx);
    if (!ent) return fail(SFB_ENOTBUILT, "step kernel for this L not compiled");
    SfbStepParams P;
    P.nlm_in = reinterpret_cast<const double2*>(nlm_in);
    P.nlm_out = reinterpret_cast<double2*>(nlm_out);
    P.ugrad = ugrad;
    P.tau = ddrx ? tau : nullptr;
    P.gamma0_arr = o->gamma0_arr;
    P.lambda_arr = (o->terms & SFB_CDRX) ? o->lambda_arr : nullptr;
    P.N = N; P.ld_in = ld_in; P.ld_out = ld_out; P.ld_u = ld_u; P.ld_t = ld_t;
    P.dt = o->dt; P.iota = o->iota; P.zeta = o->zeta; P.nu_mult = o->nu_mult;
    P.gamma0 = o->gamma0;
    P.lambda = (o->terms & SFB_CDRX) ? o->lambda : 0.0;
    P.nstage = o->scheme == SFB_RK4 ? 4 : 1;
    P.use_lrot = (o->terms & SFB_LROT) ? 1 : 0;
    P.use_reg = (o->terms & SFB_REG) ? 1 : 0;
    for (int s = 0; s < o->nsteps; ++s) {
        cudaError_t e = ent->fn(P, g.reg, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "step kernel launch");
        P.nlm_in = P.nlm_out;   // subsequent sub-steps run in place
        P.ld_in = P.ld_out;
    }
    return SFB_OK;
}

static int ensure_slot(Slot& s, size_t bytes_n, size_t nodes) {
    if (!s.st) CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    if (s.cap_n < bytes_n) {
        cudaFree(s.nin); cudaFree(s.nout);
        s.nin = s.nout = nullptr; s.cap_n = 0;
        CK(cudaMalloc(&s.nin, bytes_n));
        CK(cudaMalloc(&s.nout, bytes_n));
        s.cap_n = bytes_n;
    }
    if (s.cap_c < nodes) {
        cudaFree(s.ug); cudaFree(s.tau); cudaFree(s.g0); cudaFree(s.lam);
        s.ug = s.tau = s.g0 = s.lam = nullptr; s.cap_c = 0;
        CK(cudaMalloc(&s.ug, nodes * 9 * sizeof(double)));
        CK(cudaMalloc(&s.tau, nodes * 9 * sizeof(double)));
        CK(cudaMalloc(&s.g0, nodes * sizeof(double)));
        CK(cudaMalloc(&s.lam, nodes * sizeof(double)));
        s.cap_c = nodes;
    }
    return SFB_OK;
}

// Host-pointer variant: chunks of nodes are pipelined over 3 streams (H2D | kernel | D2H overlap).
int sfb_step_arr(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld,
                 const double* ugrad, const double* tau, const sfb_step_opts* o) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    int rc = check_opts(o);
    if (rc) return rc;
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!nlm_in || !nlm_out || !ugrad) return fail(SFB_EINVAL, "null array");
    std::lock_guard<std::mutex> lk(g.mu);
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (g_stage.dev != dev) { g_stage.release(); g_stage.dev = dev; }
    const int n = g.n;
    const int64_t chunk = std::min<int64_t>(N, 1 << 16);
    const bool ddrx = (o->terms & SFB_DDRX) != 0;
    int slot = 0;
    for (int64_t p0 = 0; p0 < N; p0 += chunk, slot = (slot + 1) % Staging::kSlots) {
        const int64_t c = std::min<int64_t>(chunk, N - p0);
        Slot& s = g_stage.s[slot];
        rc = ensure_slot(s, (size_t)chunk * n * 16, (size_t)chunk);
        if (rc) return rc;
        CK(cudaMemcpy2DAsync(s.nin, chunk * 16, nlm_in + 2 * p0, ld * 16, c * 16, n, cudaMemcpyHostToDevice, s.st));
        CK(cudaMemcpy2DAsync(s.ug, chunk * 8, ugrad + p0, ld * 8, c * 8, 9, cudaMemcpyHostToDevice, s.st));
        if (ddrx && tau) CK(cudaMemcpy2DAsync(s.tau, chunk * 8, tau + p0, ld * 8, c * 8, 9, cudaMemcpyHostToDevice, s.st));
        sfb_step_opts oo = *o;
        if (o->gamma0_arr) { CK(cudaMemcpyAsync(s.g0, o->gamma0_arr + p0, c * 8, cudaMemcpyHostToDevice, s.st)); oo.gamma0_arr = s.g0; }
        if (o->lambda_arr) { CK(cudaMemcpyAsync(s.lam, o->lambda_arr + p0, c * 8, cudaMemcpyHostToDevice, s.st)); oo.lambda_arr = s.lam; }
        rc = sfb_step_arr_dev(s.nin, s.nout, c, chunk, chunk, s.ug, chunk, (ddrx && tau) ? s.tau : nullptr, chunk, &oo, s.st);
        if (rc) return rc;
        CK(cudaMemcpy2DAsync(nlm_out + 2 * p0, ld * 16, s.nout, chunk * 16, c * 16, n, cudaMemcpyDeviceToHost, s.st));
    }
    for (auto& s : g_stage.s)
        if (s.st) CK(cudaStreamSynchronize(s.st));
    return SFB_OK;
}

int sfb_dev_malloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return fail(SFB_EINVAL, "bad args");
    CK(cudaMalloc(p, (size_t)bytes));
    return SFB_OK;
}
int sfb_dev_free(void* p) { CK(cudaFree(p)); return SFB_OK; }
int sfb_memcpy_h2d(void* dst, const void* src, int64_t bytes) { CK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice)); return SFB_OK; }
int sfb_memcpy_d2h(void* dst, const void* src, int64_t bytes) { CK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost)); return SFB_OK; }
int sfb_host_alloc_pinned(void** p, int64_t bytes) {
    if (!p || bytes < 0) return fail(SFB_EINVAL, "bad args");
    CK(cudaMallocHost(p, (size_t)bytes));
    return SFB_OK;
}
int sfb_host_free_pinned(void* p) { CK(cudaFreeHost(p)); return SFB_OK; }
int sfb_sync(void) { CK(cudaDeviceSynchronize()); return SFB_OK; }
int sfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int sfb_set_device(int dev) { CK(cudaSetDevice(dev)); return SFB_OK; }

}  // extern "C"
