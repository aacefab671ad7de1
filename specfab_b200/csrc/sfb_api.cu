// C ABI of libspecfab_b200.so (declared in include/specfab_b200.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sfb_common.cuh"
#include "sfb_fields.cuh"
#include "specfab_b200.h"

cudaError_t sfb_launch_mexport(int mode, int L, const double* a33, const double* b33, const double2* nlm, long long ldn,
                               long long N, long long ld, double iota, double zeta, double2* M, cudaStream_t st);
cudaError_t sfb_launch_mreg(int L, const SfbRegConst& reg, const double* eps, long long N, long long ld, double* M, cudaStream_t st);
void sfb_ops_release();
cudaError_t sfb_launch_bounds(const double2* in, double2* out, long long N, long long ldi, long long ldo, int n, cudaStream_t st);
cudaError_t sfb_launch_reduced(int to_reduced, const double2* src, double2* dst, long long N, long long lds, long long ldd, int L,
                               cudaStream_t st);
cudaError_t sfb_launch_bounds_rnlm(const double2* in, double2* out, long long N, long long ldi, long long ldo, int r, cudaStream_t st);
cudaError_t sfb_launch_a2(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st, int red = 0);
cudaError_t sfb_launch_a4(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st);
cudaError_t sfb_launch_eig(const double2* nlm, const double* M, long long N, long long ld, int plane, double* ei, double* lami,
                           long long ldo, cudaStream_t st, int red = 0);
cudaError_t sfb_launch_a6(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st);
cudaError_t sfb_launch_eij3(const double2* nlm, long long N, long long ld, const double* e1, const double* e2, const double* e3,
                            long long lde, const sfb::EijCoef& K, double* Eij, long long ldo, double* ei_out, double* lam_out,
                            int* status, cudaStream_t st);
cudaError_t sfb_launch_caffe(const double2* nlm, long long N, long long ld, const double* eps, long long lde, double Emin, double Emax,
                             int n_grain, double* E, cudaStream_t st);
cudaError_t sfb_launch_pfj(const double2* nlm, long long N, long long ld, int Lmax, double* J, cudaStream_t st);
cudaError_t sfb_launch_mexport_reduced(int mode, int L, const double* a33, const double* b33, const double2* nlm, long long ldn,
                                       long long N, long long ld, double iota, double zeta, double* Mrr, double* Mri, double* Mir,
                                       double* Mii, cudaStream_t st);
cudaError_t sfb_launch_reduce_dense(const double* M, int is_complex, int L, long long N, long long ldi, double* Mrr, double* Mri,
                                    double* Mir, double* Mii, cudaStream_t st);
cudaError_t sfb_launch_ingest(int rank, const double* A, long long N, long long ld, double2* out, long long ldo, cudaStream_t st);
cudaError_t sfb_launch_eij_orth(const double2* q1, long long ld1, const double2* q2, long long ld2, const double2* q3, long long ld3,
                                long long N, const double* e1, const double* e2, const double* e3, long long lde,
                                const double* Eij_grain, int n_grain, double* Eij, long long ldo, cudaStream_t st);
cudaError_t sfb_launch_eij(const double2* nlm, long long N, long long ld, const double* e1, const double* e2, const double* e3,
                           long long lde, const sfb::EijCoef& K, double* Eij, long long ldo, double* ei_out, double* lam_out,
                           int* status, cudaStream_t st, int red = 0);
cudaError_t sfb_launch_evw(const double2* nlm, long long N, long long ld, const double* v, const double* w, const double* tau, long long lde,
                           const sfb::EijCoef& K, double* Evw, int* status, cudaStream_t st);

// ---- queues of general-state tiles (sfb_common.cuh) ----
namespace {
struct WorkList { int* buf = nullptr; long long cap = 0; };
std::mutex g_wl_mu;
std::map<std::pair<int, cudaStream_t>, WorkList> g_wl;
}  // namespace
cudaError_t sfb_worklist_get(cudaStream_t st, long long ntile, int** buf) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_wl_mu);
    WorkList& w = g_wl[std::make_pair(dev, st)];
    if (w.cap < ntile) {
        if (w.buf) {                       // kernels of earlier launches on this stream may still use the old queue
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            cudaFree(w.buf);
            w.buf = nullptr; w.cap = 0;
        }
        long long cap = 4096;
        while (cap < ntile) cap *= 2;
        e = cudaMalloc(&w.buf, (size_t)cap * sizeof(int));
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(w.buf, 0, (size_t)cap * sizeof(int), st);
        if (e != cudaSuccess) return e;
        w.cap = cap;
    }
    *buf = w.buf;
    return cudaSuccess;
}
void sfb_worklist_release() {
    std::lock_guard<std::mutex> lk(g_wl_mu);
    for (auto& kv : g_wl) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(kv.first.first);
        cudaFree(kv.second.buf);
        cudaSetDevice(cur);
    }
    g_wl.clear();
}

struct SfbStepEntry {
    int L, ddrx, variant, R, TN, dfma_node;
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

// Error text is per thread (sfb_last_error reports the calling thread's last failure); the process-wide copy of the most
// recent message -- what a caller on another thread falls back to -- is only touched under its own lock.
thread_local std::string t_err;
std::mutex g_err_mu;

int fail(int code, const std::string& msg) {
    t_err = msg;
    std::lock_guard<std::mutex> lk(g_err_mu);
    g.err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(SFB_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int cuda_rc(cudaError_t e, const char* what) { return e == cudaSuccess ? SFB_OK : cuda_fail(e, what); }
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

int g_variant = 0;   // tuning knob (sfb_set_variant); falls back to variant 0 when absent
// variant 0 is the default; variant 100, when compiled, is the default for multi-stage (RK4) steps (build.py TUNE_RK)
const SfbStepEntry* find_step(int L, int ddrx, int nstage) {
    const SfbStepEntry *def = nullptr, *rk = nullptr;
    for (const auto& e : kStepRegistry)
        if (e.L == L && e.ddrx == ddrx) {
            if (g_variant != 0 && e.variant == g_variant) return &e;
            if (e.variant == 0) def = &e;
            if (e.variant == 100) rk = &e;
        }
    return (g_variant == 0 && nstage > 1 && rk) ? rk : def;
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
// one staging ring and one lock per device: host-pointer calls on different devices (sfb_step_arr_multi, or a caller's own
// threads after sfb_set_device) run concurrently; calls on the same device take turns
constexpr int kMaxDev = 64;
Staging g_stage[kMaxDev];
std::mutex g_stage_mu[kMaxDev];

}  // namespace

extern "C" {

const char* sfb_last_error(void) {
    if (t_err.empty()) {          // nothing failed on this thread: hand out a private copy of the process-wide message
        std::lock_guard<std::mutex> lk(g_err_mu);
        t_err = g.err;
    }
    return t_err.c_str();
}

int sfb_init(int L) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (L % 2 != 0 || L < 4 || L > SFB_MAXL) return fail(SFB_EINVAL, "sfb_init: L must be even, 4 <= L <= 20");
    if (!find_step(L, 0, 1) || !find_step(L, 1, 1)) return fail(SFB_ENOTBUILT, "sfb_init: kernels for this L were not compiled");
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
    {
        int cur = 0;
        cudaGetDevice(&cur);
        for (int d = 0; d < kMaxDev; ++d) {
            std::lock_guard<std::mutex> sl(g_stage_mu[d]);
            if (g_stage[d].dev >= 0) { cudaSetDevice(d); g_stage[d].release(); }
        }
        cudaSetDevice(cur);
    }
    sfb_ops_release();
    sfb_worklist_release();
    g.L = 0; g.n = 0;
}

const char* sfb_build_info(void) {
    if (g.info.empty()) {
        std::string s = "{\"arch\":\"sm_100a\",\"step_kernels\":[";
        bool first = true;
        for (const auto& e : kStepRegistry) {
            char b[160];
            snprintf(b, sizeof b, "%s{\"L\":%d,\"ddrx\":%d,\"variant\":%d,\"roles\":%d,\"tile\":%d,\"dfma_per_node_rhs\":%d}", first ? "" : ",", e.L, e.ddrx, e.variant, e.R, e.TN, e.dfma_node);
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
    if (o->reserved & ~SFB_STEP_GENERAL) return fail(SFB_EINVAL, "unknown flag in sfb_step_opts.reserved");
    return SFB_OK;
}

// rio = 1: the state arrays are in reduced form (rows m >= 0 only, sfb_step_rnlm_arr*)
static int step_dev_impl(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                         const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                         const sfb_step_opts* o, void* stream, int rio) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    int rc = check_opts(o);
    if (rc) return rc;
    if (N < 0 || ld_in < N || ld_out < N || ld_u < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!nlm_in || !nlm_out || !ugrad) return fail(SFB_EINVAL, "null array");
    if (((uintptr_t)nlm_in & 15) || ((uintptr_t)nlm_out & 15)) return fail(SFB_EINVAL, "nlm arrays must be 16-byte aligned");
    const int ddrx = (o->terms & SFB_DDRX) ? 1 : 0;
    if (ddrx && tau && ld_t < N) return fail(SFB_EINVAL, "ld_t < N");
    const SfbStepEntry* ent = find_step(g.L, ddrx, o->scheme == SFB_RK4 ? 4 : 1);
    if (o->reserved & SFB_STEP_GENERAL) {      // every tile through the general (unreduced) algorithm: the full-form kernels
        if (rio) return fail(SFB_EINVAL, "SFB_STEP_GENERAL: reduced-form states are real ODFs by construction");
        ent = nullptr;
        for (const auto& e : kStepRegistry)
            if (e.L == g.L && e.ddrx == ddrx && e.variant == 40) ent = &e;
    }
    if (!ent) return fail(SFB_ENOTBUILT, "step kernel for this L not compiled");
    SfbStepParams P;
    P.rio = rio;
    P.ktab = nullptr; P.raw_ok = 0; P.n0_global = 0;
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
        if (e == cudaErrorNotSupported) return fail(SFB_EINVAL, "the selected kernel variant (sfb_set_variant) has no reduced-form I/O");
        if (e != cudaSuccess) return cuda_fail(e, "step kernel launch");
        P.nlm_in = P.nlm_out;   // subsequent sub-steps run in place
        P.ld_in = P.ld_out;
    }
    return SFB_OK;
}

int sfb_step_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                     const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                     const sfb_step_opts* o, void* stream) {
    return step_dev_impl(nlm_in, nlm_out, N, ld_in, ld_out, ugrad, ld_u, tau, ld_t, o, stream, 0);
}

int sfb_step_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                          const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                          const sfb_step_opts* o, void* stream) {
    return step_dev_impl(rnlm_in, rnlm_out, N, ld_in, ld_out, ugrad, ld_u, tau, ld_t, o, stream, 1);
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
static int step_host_impl(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld,
                          const double* ugrad, const double* tau, const sfb_step_opts* o, int rio) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    int rc = check_opts(o);
    if (rc) return rc;
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!nlm_in || !nlm_out || !ugrad) return fail(SFB_EINVAL, "null array");
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return fail(SFB_EINVAL, "device ordinal out of range");
    std::lock_guard<std::mutex> lk(g_stage_mu[dev]);
    Staging& stage = g_stage[dev];
    stage.dev = dev;
    const int n = rio ? (g.L / 2 + 1) * (g.L / 2 + 1) : g.n;      // coefficient rows that cross PCIe
    static const int64_t chunk_nodes = [] {      // nodes per pipeline stage (H2D | kernel | D2H on rotating streams); SFB_CHUNK overrides
        const char* ev = getenv("SFB_CHUNK");
        const int64_t c = ev ? atoll(ev) : (1 << 16);
        return c < 1024 ? (int64_t)1024 : c;
    }();
    const int64_t chunk = std::min<int64_t>(N, chunk_nodes);
    const bool ddrx = (o->terms & SFB_DDRX) != 0;
    int slot = 0;
    rc = SFB_OK;
    for (int64_t p0 = 0; p0 < N && rc == SFB_OK; p0 += chunk, slot = (slot + 1) % Staging::kSlots) {
        const int64_t c = std::min<int64_t>(chunk, N - p0);
        Slot& s = stage.s[slot];
        rc = ensure_slot(s, (size_t)chunk * n * 16, (size_t)chunk);
        if (rc) break;
        // a coefficient row of the chunk is one contiguous run of c nodes on both sides: n (or 9) descriptors of c*16 (c*8) bytes
        rc = cuda_rc(cudaMemcpy2DAsync(s.nin, chunk * 16, nlm_in + 2 * p0, ld * 16, c * 16, n, cudaMemcpyHostToDevice, s.st), "H2D nlm");
        if (!rc) rc = cuda_rc(cudaMemcpy2DAsync(s.ug, chunk * 8, ugrad + p0, ld * 8, c * 8, 9, cudaMemcpyHostToDevice, s.st), "H2D ugrad");
        if (!rc && ddrx && tau) rc = cuda_rc(cudaMemcpy2DAsync(s.tau, chunk * 8, tau + p0, ld * 8, c * 8, 9, cudaMemcpyHostToDevice, s.st), "H2D tau");
        sfb_step_opts oo = *o;
        if (!rc && o->gamma0_arr) { rc = cuda_rc(cudaMemcpyAsync(s.g0, o->gamma0_arr + p0, c * 8, cudaMemcpyHostToDevice, s.st), "H2D gamma0"); oo.gamma0_arr = s.g0; }
        if (!rc && o->lambda_arr) { rc = cuda_rc(cudaMemcpyAsync(s.lam, o->lambda_arr + p0, c * 8, cudaMemcpyHostToDevice, s.st), "H2D lambda"); oo.lambda_arr = s.lam; }
        if (!rc) rc = step_dev_impl(s.nin, s.nout, c, chunk, chunk, s.ug, chunk, (ddrx && tau) ? s.tau : nullptr, chunk, &oo, s.st, rio);
        if (!rc) rc = cuda_rc(cudaMemcpy2DAsync(nlm_out + 2 * p0, ld * 16, s.nout, chunk * 16, c * 16, n, cudaMemcpyDeviceToHost, s.st), "D2H nlm");
    }
    // also on the error path: the slot streams may still be reading from / writing to the caller's arrays
    for (auto& s : stage.s)
        if (s.st) {
            const cudaError_t e = cudaStreamSynchronize(s.st);
            if (e != cudaSuccess && rc == SFB_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
        }
    return rc;
}

int sfb_step_arr(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld,
                 const double* ugrad, const double* tau, const sfb_step_opts* o) {
    return step_host_impl(nlm_in, nlm_out, N, ld, ugrad, tau, o, 0);
}

int sfb_step_rnlm_arr(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld,
                      const double* ugrad, const double* tau, const sfb_step_opts* o) {
    return step_host_impl(rnlm_in, rnlm_out, N, ld, ugrad, tau, o, 1);
}

// One host array sharded over several GPUs of this process: contiguous node ranges [i*N/G, (i+1)*N/G), one host
// thread + one staging ring (three streams) per device, no inter-GPU traffic.  The reference's callers are single-process
// programs (src/dynamics.f90:99-110 via Fortran / f2py / Elmer): this is how ONE such process uses all GPUs of the box.
static int step_multi_impl(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld, const double* ugrad, const double* tau,
                           const sfb_step_opts* o, const int* devices, int ndev, int rio) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (ndev < 1 || !devices) return fail(SFB_EINVAL, "need at least one device");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    for (int i = 0; i < ndev; ++i) {
        if (devices[i] < 0 || devices[i] >= count || devices[i] >= kMaxDev) return fail(SFB_EINVAL, "device ordinal out of range");
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(SFB_EINVAL, "duplicate device in the list");
    }
    if (N == 0) return check_opts(o);
    int cur = 0;
    CK(cudaGetDevice(&cur));
    std::vector<int> rcs(ndev, SFB_OK);
    std::vector<std::string> errs(ndev);
    auto work = [&](int i) {
        // ranges on 32-node boundaries, so that the tiles -- and with them every bit of the result -- do not depend on G
        const int64_t tiles = (N + 31) / 32;
        const int64_t p0 = std::min<int64_t>(N, (tiles * i / ndev) * 32), p1 = std::min<int64_t>(N, (tiles * (i + 1) / ndev) * 32);
        if (p1 <= p0) return;
        if (cudaSetDevice(devices[i]) != cudaSuccess) { rcs[i] = SFB_ECUDA; errs[i] = "cudaSetDevice failed"; return; }
        sfb_step_opts oo = *o;
        if (o->gamma0_arr) oo.gamma0_arr = o->gamma0_arr + p0;
        if (o->lambda_arr) oo.lambda_arr = o->lambda_arr + p0;
        rcs[i] = step_host_impl(nlm_in + 2 * p0, nlm_out + 2 * p0, p1 - p0, ld, ugrad + p0, tau ? tau + p0 : nullptr, &oo, rio);
        if (rcs[i]) errs[i] = t_err;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < ndev; ++i) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    cudaSetDevice(cur);
    for (int i = 0; i < ndev; ++i)
        if (rcs[i]) return fail(rcs[i], "device " + std::to_string(devices[i]) + ": " + errs[i]);
    return SFB_OK;
}

int sfb_step_arr_multi(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld, const double* ugrad, const double* tau,
                       const sfb_step_opts* o, const int* devices, int ndev) {
    return step_multi_impl(nlm_in, nlm_out, N, ld, ugrad, tau, o, devices, ndev, 0);
}

int sfb_step_rnlm_arr_multi(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld, const double* ugrad, const double* tau,
                            const sfb_step_opts* o, const int* devices, int ndev) {
    return step_multi_impl(rnlm_in, rnlm_out, N, ld, ugrad, tau, o, devices, ndev, 1);
}

int sfb_set_variant(int v) { g_variant = v; return SFB_OK; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// a2 / a4 / eig / Eij
// ---------------------------------------------------------------------------------------------
namespace {
struct DevTmp {     // scratch device allocation of the host-pointer entry points
    void* p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
int basic_check(const void* nlm, int64_t N, int64_t ld) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N > 0 && !nlm) return fail(SFB_EINVAL, "null array");
    return SFB_OK;
}
int plane_code(const char* plane) {
    if (!plane) return -1;
    if (plane[0] == 'i' && plane[1] == 'j') return 0;
    if (plane[0] == 'x' && plane[1] == 'y') return 1;
    if (plane[0] == 'x' && plane[1] == 'z') return 2;
    return -1;
}
// src/rheologies.f90:123-135 with d = 3, evaluated on the host with libm pow (like the reference)
void rheo_params(const double E[2], double n, int ef, double& cI, double& cM, double& cL) {
    const int d = 3;
    const double ne = ef * 2 / (n + 1);
    cI = (pow(E[0], ne) - 1) / (d - 1);
    cM = (d * (pow(E[0], ne) + 1) - 2) / (d - 1) - 2 * pow(E[1], ne);
    cL = pow(E[1], ne) - 1;
}
int make_coef(const double* Eij_grain, double alpha, int n_grain, sfb::EijCoef& K) {
    if (!Eij_grain) return fail(SFB_EINVAL, "null Eij_grain");
    if (n_grain != 1 && n_grain != 3 && n_grain != -3) return fail(SFB_EINVAL, "unsupported n' (n_grain must be 1, 3 or -3)");
    if (n_grain == 3 && g.L < 8) return fail(SFB_EINVAL, "Sachs homogenization with n'=3 requires L >= 8");
    rheo_params(Eij_grain, (double)n_grain, 1, K.sA, K.sB, K.sC);
    rheo_params(Eij_grain, (double)n_grain, -1, K.tA, K.tB, K.tC);
    // n' = -3: numerator and isotropic denominator both carry I2 = tau:tau, which cancels (src/homogenizations.f90:105-109,209)
    K.s_iso = n_grain == -3 ? 1.0 : 1 + 2.0 / 15 * K.sB + 2.0 / 3 * K.sC;
    K.t_iso = 1 + 2.0 / 15 * K.tB + 2.0 / 3 * K.tC;
    K.alpha = alpha;
    return SFB_OK;
}
// copy the first `rows` coefficient rows of a host nlm(N,n) to a packed device array [rows][N]
int stage_rows(DevTmp& d, const double* nlm, int64_t N, int64_t ld, int rows) {
    CK(d.alloc((size_t)rows * N * 16));
    CK(cudaMemcpy2D(d.p, (size_t)N * 16, nlm, (size_t)ld * 16, (size_t)N * 16, rows, cudaMemcpyHostToDevice));
    return SFB_OK;
}
}  // namespace

extern "C" {

int sfb_a2_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a2, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (N && !a2) return fail(SFB_EINVAL, "null output");
    CK(sfb_launch_a2(reinterpret_cast<const double2*>(nlm), N, ld, a2, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_a2_arr(const double* nlm, int64_t N, int64_t ld, double* a2) {
    int rc = basic_check(nlm, N, ld);
    if (rc || N == 0) return rc;
    DevTmp in, out;
    if ((rc = stage_rows(in, nlm, N, ld, 6))) return rc;
    CK(out.alloc((size_t)N * 9 * 8));
    if ((rc = sfb_a2_arr_dev(in.as<double>(), N, N, out.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(a2, out.p, (size_t)N * 9 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_a4_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a4, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (N && !a4) return fail(SFB_EINVAL, "null output");
    CK(sfb_launch_a4(reinterpret_cast<const double2*>(nlm), N, ld, a4, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_a4_arr(const double* nlm, int64_t N, int64_t ld, double* a4) {
    int rc = basic_check(nlm, N, ld);
    if (rc || N == 0) return rc;
    DevTmp in, out;
    if ((rc = stage_rows(in, nlm, N, ld, 15))) return rc;
    CK(out.alloc((size_t)N * 81 * 8));
    if ((rc = sfb_a4_arr_dev(in.as<double>(), N, N, out.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(a4, out.p, (size_t)N * 81 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_eig_arr_dev(const double* nlm, int64_t N, int64_t ld, double* ei, double* lami, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (N && (!ei || !lami)) return fail(SFB_EINVAL, "null output");
    CK(sfb_launch_eig(reinterpret_cast<const double2*>(nlm), nullptr, N, ld, 0, ei, lami, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_eig_arr(const double* nlm, int64_t N, int64_t ld, double* ei, double* lami) {
    int rc = basic_check(nlm, N, ld);
    if (rc || N == 0) return rc;
    DevTmp in, o1, o2;
    if ((rc = stage_rows(in, nlm, N, ld, 6))) return rc;
    CK(o1.alloc((size_t)N * 9 * 8));
    CK(o2.alloc((size_t)N * 3 * 8));
    if ((rc = sfb_eig_arr_dev(in.as<double>(), N, N, o1.as<double>(), o2.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(ei, o1.p, (size_t)N * 9 * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lami, o2.p, (size_t)N * 3 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_eigframe_arr_dev(const double* M, int64_t N, int64_t ld, const char* plane, double* ei, double* lami, void* stream) {
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    const int pc = plane_code(plane);
    if (pc < 0) return fail(SFB_EINVAL, "eigframe: argument \"plane\" was not any of ij,xy,xz");   // src/frames.f90:48
    if (N == 0) return SFB_OK;
    if (!M || !ei || !lami) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_eig(nullptr, M, N, ld, pc, ei, lami, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_eigframe_arr(const double* M, int64_t N, const char* plane, double* ei, double* lami) {
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (plane_code(plane) < 0) return fail(SFB_EINVAL, "eigframe: argument \"plane\" was not any of ij,xy,xz");
    if (N == 0) return SFB_OK;
    if (!M || !ei || !lami) return fail(SFB_EINVAL, "null array");
    DevTmp in, o1, o2;
    CK(in.alloc((size_t)N * 9 * 8));
    CK(cudaMemcpy(in.p, M, (size_t)N * 9 * 8, cudaMemcpyHostToDevice));
    CK(o1.alloc((size_t)N * 9 * 8));
    CK(o2.alloc((size_t)N * 3 * 8));
    int rc = sfb_eigframe_arr_dev(in.as<double>(), N, N, plane, o1.as<double>(), o2.as<double>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(ei, o1.p, (size_t)N * 9 * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lami, o2.p, (size_t)N * 3 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_Eij_tranisotropic_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* e1, const double* e2, const double* e3,
                                  const double* Eij_grain, double alpha, int n_grain, double* Eij, int32_t* status, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    sfb::EijCoef K;
    if ((rc = make_coef(Eij_grain, alpha, n_grain, K))) return rc;
    if (N == 0) return SFB_OK;
    if (!e1 || !e2 || !e3 || !Eij) return fail(SFB_EINVAL, "null array");
    if (n_grain == 3)
        CK(sfb_launch_eij3(reinterpret_cast<const double2*>(nlm), N, ld, e1, e2, e3, N, K, Eij, N, nullptr, nullptr, status, (cudaStream_t)stream));
    else
        CK(sfb_launch_eij(reinterpret_cast<const double2*>(nlm), N, ld, e1, e2, e3, N, K, Eij, N, nullptr, nullptr, status, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_Eij_tranisotropic_arr(const double* nlm, int64_t N, int64_t ld, const double* e1, const double* e2, const double* e3,
                              const double* Eij_grain, double alpha, int n_grain, double* Eij, int32_t* status) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    sfb::EijCoef K;
    if ((rc = make_coef(Eij_grain, alpha, n_grain, K))) return rc;
    if (N == 0) return SFB_OK;
    if (!e1 || !e2 || !e3 || !Eij) return fail(SFB_EINVAL, "null array");
    DevTmp in, de, out, ds;
    if ((rc = stage_rows(in, nlm, N, ld, n_grain == 3 ? 45 : 15))) return rc;
    CK(de.alloc((size_t)N * 9 * 8));
    CK(cudaMemcpy(de.as<double>(), e1, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(de.as<double>() + 3 * N, e2, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(de.as<double>() + 6 * N, e3, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(out.alloc((size_t)N * 6 * 8));
    CK(ds.alloc((size_t)N * 4));
    rc = sfb_Eij_tranisotropic_arr_dev(in.as<double>(), N, N, de.as<double>(), de.as<double>() + 3 * N, de.as<double>() + 6 * N,
                                       Eij_grain, alpha, n_grain, out.as<double>(), ds.as<int32_t>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(Eij, out.p, (size_t)N * 6 * 8, cudaMemcpyDeviceToHost));
    if (status) CK(cudaMemcpy(status, ds.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_Evw_tranisotropic_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* v, const double* w, const double* tau,
                                  const double* Eij_grain, double alpha, int n_grain, double* Evw, int32_t* status, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    sfb::EijCoef K;
    if ((rc = make_coef(Eij_grain, alpha, n_grain, K))) return rc;
    if (n_grain == 3) return fail(SFB_EINVAL, "Evw_tranisotropic_arr: n_grain = 3 is served by Eij_tranisotropic_arr only (n' = 1, -3 here)");
    if (N == 0) return SFB_OK;
    if (!v || !w || !tau || !Evw) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_evw(reinterpret_cast<const double2*>(nlm), N, ld, v, w, tau, N, K, Evw, status, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_Evw_tranisotropic_arr(const double* nlm, int64_t N, int64_t ld, const double* v, const double* w, const double* tau,
                              const double* Eij_grain, double alpha, int n_grain, double* Evw, int32_t* status) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (N == 0) { sfb::EijCoef K; return make_coef(Eij_grain, alpha, n_grain, K); }
    if (!v || !w || !tau || !Evw) return fail(SFB_EINVAL, "null array");
    DevTmp in, dv, out, ds;
    if ((rc = stage_rows(in, nlm, N, ld, 15))) return rc;
    CK(dv.alloc((size_t)N * 15 * 8));
    CK(cudaMemcpy(dv.as<double>(), v, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv.as<double>() + 3 * N, w, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dv.as<double>() + 6 * N, tau, (size_t)N * 9 * 8, cudaMemcpyHostToDevice));
    CK(out.alloc((size_t)N * 8));
    CK(ds.alloc((size_t)N * 4));
    rc = sfb_Evw_tranisotropic_arr_dev(in.as<double>(), N, N, dv.as<double>(), dv.as<double>() + 3 * N, dv.as<double>() + 6 * N,
                                       Eij_grain, alpha, n_grain, out.as<double>(), ds.as<int32_t>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(Evw, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    if (status) CK(cudaMemcpy(status, ds.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_step_moments_Eij_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                                 const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t, const sfb_step_opts* opts,
                                 const double* Eij_grain, double alpha, int n_grain,
                                 double* Eij, double* a2, double* a4, double* ei, double* lami, int32_t* status, void* stream) {
    int rc = sfb_step_arr_dev(nlm_in, nlm_out, N, ld_in, ld_out, ugrad, ld_u, tau, ld_t, opts, stream);
    if (rc) return rc;
    if (a2 && (rc = sfb_a2_arr_dev(nlm_out, N, ld_out, a2, stream))) return rc;
    if (a4 && (rc = sfb_a4_arr_dev(nlm_out, N, ld_out, a4, stream))) return rc;
    return sfb_Eij_eigenframe_arr_dev(nlm_out, N, ld_out, Eij_grain, alpha, n_grain, Eij, ei, lami, status, stream);
}
int sfb_a6_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a6, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (g.L < 6) return fail(SFB_EINVAL, "a6 requires L >= 6");
    if (N && !a6) return fail(SFB_EINVAL, "null output");
    CK(sfb_launch_a6(reinterpret_cast<const double2*>(nlm), N, ld, a6, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_a6_arr(const double* nlm, int64_t N, int64_t ld, double* a6) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (g.L < 6) return fail(SFB_EINVAL, "a6 requires L >= 6");
    if (N == 0) return SFB_OK;
    DevTmp in, out;
    if ((rc = stage_rows(in, nlm, N, ld, 28))) return rc;
    CK(out.alloc((size_t)N * 729 * 8));
    if ((rc = sfb_a6_arr_dev(in.as<double>(), N, N, out.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(a6, out.p, (size_t)N * 729 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_E_CAFFE_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* eps, double Emin, double Emax, int n_grain, double* E,
                        void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (n_grain == 3 && g.L < 8) return fail(SFB_EINVAL, "E_CAFFE with n'=3 requires L >= 8");
    if (N && (!eps || !E)) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_caffe(reinterpret_cast<const double2*>(nlm), N, ld, eps, N, Emin, Emax, n_grain, E, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_E_CAFFE_arr(const double* nlm, int64_t N, int64_t ld, const double* eps, double Emin, double Emax, int n_grain, double* E) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (n_grain == 3 && g.L < 8) return fail(SFB_EINVAL, "E_CAFFE with n'=3 requires L >= 8");
    if (N == 0) return SFB_OK;
    if (!eps || !E) return fail(SFB_EINVAL, "null array");
    DevTmp in, de, out;
    if ((rc = stage_rows(in, nlm, N, ld, n_grain == 3 ? 45 : 15))) return rc;
    CK(de.alloc((size_t)N * 9 * 8));
    CK(cudaMemcpy(de.p, eps, (size_t)N * 9 * 8, cudaMemcpyHostToDevice));
    CK(out.alloc((size_t)N * 8));
    if ((rc = sfb_E_CAFFE_arr_dev(in.as<double>(), N, N, de.as<double>(), Emin, Emax, n_grain, out.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(E, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_pfJ_arr_dev(const double* nlm, int64_t N, int64_t ld, int Lmax, double* J, void* stream) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (Lmax < 0 || Lmax > g.L || (Lmax & 1)) return fail(SFB_EINVAL, "need even 0 <= Lmax <= L");
    if (N && !J) return fail(SFB_EINVAL, "null output");
    CK(sfb_launch_pfj(reinterpret_cast<const double2*>(nlm), N, ld, Lmax, J, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_pfJ_arr(const double* nlm, int64_t N, int64_t ld, int Lmax, double* J) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    if (Lmax < 0 || Lmax > g.L || (Lmax & 1)) return fail(SFB_EINVAL, "need even 0 <= Lmax <= L");
    if (N == 0) return SFB_OK;
    DevTmp in, out;
    const int rows = (Lmax + 1) * (Lmax + 2) / 2;
    if ((rc = stage_rows(in, nlm, N, ld, rows))) return rc;
    CK(out.alloc((size_t)N * 8));
    if ((rc = sfb_pfJ_arr_dev(in.as<double>(), N, N, Lmax, out.as<double>(), nullptr))) return rc;
    CK(cudaMemcpy(J, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_ai_to_nlm_arr_dev(int rank, const double* a, int64_t N, int64_t ld, double* nlm, int64_t ld_nlm, void* stream) {
    if (rank != 2 && rank != 4 && rank != 6) return fail(SFB_EINVAL, "rank must be 2, 4 or 6");
    if (N < 0 || ld < N || ld_nlm < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!a || !nlm) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_ingest(rank, a, N, ld, reinterpret_cast<double2*>(nlm), ld_nlm, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_ai_to_nlm_arr(int rank, const double* a, int64_t N, double* nlm) {
    if (rank != 2 && rank != 4 && rank != 6) return fail(SFB_EINVAL, "rank must be 2, 4 or 6");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!a || !nlm) return fail(SFB_EINVAL, "null array");
    const int ne = rank == 2 ? 9 : (rank == 4 ? 81 : 729), nrow = (rank + 1) * (rank + 2) / 2;
    DevTmp di, dout;
    CK(di.alloc((size_t)N * ne * 8));
    CK(cudaMemcpy(di.p, a, (size_t)N * ne * 8, cudaMemcpyHostToDevice));
    CK(dout.alloc((size_t)N * nrow * 16));
    int rc = sfb_ai_to_nlm_arr_dev(rank, di.as<double>(), N, N, dout.as<double>(), N, nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(nlm, dout.p, (size_t)N * nrow * 16, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_Eij_orthotropic_arr_dev(const double* nlm_1, int64_t ld1, const double* nlm_2, int64_t ld2, const double* nlm_3, int64_t ld3,
                                int64_t N, const double* e1, const double* e2, const double* e3,
                                const double* Eij_grain, double alpha, int n_grain, double* Eij, void* stream) {
    (void)alpha;   // Sachs only, src/enhancementfactors.f90:160-162
    int rc = basic_check(nlm_1, N, ld1);
    if (rc) return rc;
    if ((rc = basic_check(nlm_2, N, ld2))) return rc;
    if (nlm_3 && ld3 < N) return fail(SFB_EINVAL, "need N <= ld3");
    if (!Eij_grain) return fail(SFB_EINVAL, "null Eij_grain");
    if (N == 0) return SFB_OK;
    if (!e1 || !e2 || !e3 || !Eij) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_eij_orth(reinterpret_cast<const double2*>(nlm_1), ld1, reinterpret_cast<const double2*>(nlm_2), ld2,
                           reinterpret_cast<const double2*>(nlm_3), ld3, N, e1, e2, e3, N, Eij_grain, n_grain, Eij, N, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_Eij_orthotropic_arr(const double* nlm_1, const double* nlm_2, const double* nlm_3, int64_t N, int64_t ld,
                            const double* e1, const double* e2, const double* e3,
                            const double* Eij_grain, double alpha, int n_grain, double* Eij) {
    int rc = basic_check(nlm_1, N, ld);
    if (rc) return rc;
    if ((rc = basic_check(nlm_2, N, ld))) return rc;
    if (!Eij_grain) return fail(SFB_EINVAL, "null Eij_grain");
    if (N == 0) return SFB_OK;
    if (!e1 || !e2 || !e3 || !Eij) return fail(SFB_EINVAL, "null array");
    DevTmp q1, q2, q3, de, out;
    if ((rc = stage_rows(q1, nlm_1, N, ld, 15))) return rc;
    if ((rc = stage_rows(q2, nlm_2, N, ld, 15))) return rc;
    if (nlm_3 && (rc = stage_rows(q3, nlm_3, N, ld, 15))) return rc;
    CK(de.alloc((size_t)N * 9 * 8));
    CK(cudaMemcpy(de.as<double>(), e1, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(de.as<double>() + 3 * N, e2, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(de.as<double>() + 6 * N, e3, (size_t)N * 3 * 8, cudaMemcpyHostToDevice));
    CK(out.alloc((size_t)N * 6 * 8));
    rc = sfb_Eij_orthotropic_arr_dev(q1.as<double>(), N, q2.as<double>(), N, nlm_3 ? q3.as<double>() : nullptr, N, N,
                                     de.as<double>(), de.as<double>() + 3 * N, de.as<double>() + 6 * N, Eij_grain, alpha, n_grain,
                                     out.as<double>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(Eij, out.p, (size_t)N * 6 * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
// red = 1: the state array is in reduced form (rows m >= 0; n' = 1 and -3 only: the n' = 3 closure reads rows up to l = 8)
static int eij_eigenframe_dev(const double* nlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                              double* Eij, double* ei, double* lami, int32_t* status, void* stream, int red) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    sfb::EijCoef K;
    if ((rc = make_coef(Eij_grain, alpha, n_grain, K))) return rc;
    if (red && n_grain == 3) return fail(SFB_EINVAL, "reduced-form states: n_grain must be 1 or -3");
    if (N == 0) return SFB_OK;
    if (!Eij) return fail(SFB_EINVAL, "null array");
    if ((ei == nullptr) != (lami == nullptr)) return fail(SFB_EINVAL, "ei and lami must both be given or both be NULL");
    if (n_grain == 3)
        CK(sfb_launch_eij3(reinterpret_cast<const double2*>(nlm), N, ld, nullptr, nullptr, nullptr, 0, K, Eij, N, ei, lami, status, (cudaStream_t)stream));
    else
        CK(sfb_launch_eij(reinterpret_cast<const double2*>(nlm), N, ld, nullptr, nullptr, nullptr, 0, K, Eij, N, ei, lami, status, (cudaStream_t)stream, red));
    return SFB_OK;
}
int sfb_Eij_eigenframe_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                               double* Eij, double* ei, double* lami, int32_t* status, void* stream) {
    return eij_eigenframe_dev(nlm, N, ld, Eij_grain, alpha, n_grain, Eij, ei, lami, status, stream, 0);
}
int sfb_Eij_eigenframe_rnlm_arr_dev(const double* rnlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                                    double* Eij, double* a2, double* ei, double* lami, int32_t* status, void* stream) {
    int rc = basic_check(rnlm, N, ld);
    if (rc) return rc;
    if (a2 && N > 0) CK(sfb_launch_a2(reinterpret_cast<const double2*>(rnlm), N, ld, a2, N, (cudaStream_t)stream, 1));
    return eij_eigenframe_dev(rnlm, N, ld, Eij_grain, alpha, n_grain, Eij, ei, lami, status, stream, 1);
}
int sfb_step_moments_Eij_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                                      const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t, const sfb_step_opts* opts,
                                      const double* Eij_grain, double alpha, int n_grain,
                                      double* Eij, double* a2, double* ei, double* lami, int32_t* status, void* stream) {
    int rc = sfb_step_rnlm_arr_dev(rnlm_in, rnlm_out, N, ld_in, ld_out, ugrad, ld_u, tau, ld_t, opts, stream);
    if (rc) return rc;
    return sfb_Eij_eigenframe_rnlm_arr_dev(rnlm_out, N, ld_out, Eij_grain, alpha, n_grain, Eij, a2, ei, lami, status, stream);
}
int sfb_Eij_eigenframe_arr(const double* nlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                           double* Eij, double* ei, double* lami, int32_t* status) {
    int rc = basic_check(nlm, N, ld);
    if (rc) return rc;
    sfb::EijCoef K;
    if ((rc = make_coef(Eij_grain, alpha, n_grain, K))) return rc;
    if (N == 0) return SFB_OK;
    if (!Eij) return fail(SFB_EINVAL, "null array");
    if ((ei == nullptr) != (lami == nullptr)) return fail(SFB_EINVAL, "ei and lami must both be given or both be NULL");
    DevTmp in, out, o1, o2, ds;
    if ((rc = stage_rows(in, nlm, N, ld, n_grain == 3 ? 45 : 15))) return rc;
    CK(out.alloc((size_t)N * 6 * 8));
    CK(o1.alloc((size_t)N * 9 * 8));
    CK(o2.alloc((size_t)N * 3 * 8));
    CK(ds.alloc((size_t)N * 4));
    rc = sfb_Eij_eigenframe_arr_dev(in.as<double>(), N, N, Eij_grain, alpha, n_grain, out.as<double>(), o1.as<double>(), o2.as<double>(),
                                    ds.as<int32_t>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(Eij, out.p, (size_t)N * 6 * 8, cudaMemcpyDeviceToHost));
    if (ei) {
        CK(cudaMemcpy(ei, o1.p, (size_t)N * 9 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(lami, o2.p, (size_t)N * 3 * 8, cudaMemcpyDeviceToHost));
    }
    if (status) CK(cudaMemcpy(status, ds.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    return SFB_OK;
}

// ---------------------------------------------------------------------------------------------
// operator export: dense (N, n, n) operators, Fortran order (node contiguous)
// ---------------------------------------------------------------------------------------------
static int mexport_dev(int mode, const double* a33, const double* b33, const double* nlm, int64_t ldn, int64_t N, int64_t ld,
                       double iota, double zeta, double* M, void* stream) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!a33 || !M || (mode == 0 && !b33) || (mode == 2 && (!nlm || ldn < N))) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_mexport(mode, g.L, a33, b33, reinterpret_cast<const double2*>(nlm), ldn, N, ld, iota, zeta,
                          reinterpret_cast<double2*>(M), (cudaStream_t)stream));
    return SFB_OK;
}
static int mexport_host(int mode, const double* a33, const double* b33, const double* nlm, int64_t ldn, int64_t N,
                        double iota, double zeta, double* M) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!a33 || !M || (mode == 0 && !b33) || (mode == 2 && !nlm)) return fail(SFB_EINVAL, "null array");
    DevTmp da, db, dn, dm;
    int rc;
    CK(da.alloc((size_t)N * 72));
    CK(cudaMemcpy(da.p, a33, (size_t)N * 72, cudaMemcpyHostToDevice));
    if (mode == 0) { CK(db.alloc((size_t)N * 72)); CK(cudaMemcpy(db.p, b33, (size_t)N * 72, cudaMemcpyHostToDevice)); }
    if (mode == 2 && (rc = stage_rows(dn, nlm, N, ldn, 15))) return rc;
    const size_t mbytes = (size_t)N * g.n * g.n * 16;
    CK(dm.alloc(mbytes));
    rc = mexport_dev(mode, da.as<double>(), db.as<double>(), dn.as<double>(), N, N, N, iota, zeta, dm.as<double>(), nullptr);
    if (rc) return rc;
    CK(cudaMemcpy(M, dm.p, mbytes, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_M_LROT_arr(const double* eps, const double* omg, int64_t N, double iota, double zeta, double* M) {
    return mexport_host(0, eps, omg, nullptr, 0, N, iota, zeta, M);
}
int sfb_M_LROT_arr_dev(const double* eps, const double* omg, int64_t N, int64_t ld, double iota, double zeta, double* M, void* stream) {
    return mexport_dev(0, eps, omg, nullptr, 0, N, ld, iota, zeta, M, stream);
}
int sfb_M_DDRX_src_arr(const double* tau, int64_t N, double* M) { return mexport_host(1, tau, nullptr, nullptr, 0, N, 0, 0, M); }
int sfb_M_DDRX_src_arr_dev(const double* tau, int64_t N, int64_t ld, double* M, void* stream) {
    return mexport_dev(1, tau, nullptr, nullptr, 0, N, ld, 0, 0, M, stream);
}
int sfb_M_DDRX_arr(const double* nlm, int64_t ldn, const double* tau, int64_t N, double* M) {
    return mexport_host(2, tau, nullptr, nlm, ldn, N, 0, 0, M);
}
int sfb_M_DDRX_arr_dev(const double* nlm, int64_t ldn, const double* tau, int64_t N, int64_t ld, double* M, void* stream) {
    return mexport_dev(2, tau, nullptr, nlm, ldn, N, ld, 0, 0, M, stream);
}
// ---- reduced-form operators (src/reducedform.f90:76-120)
static int mreduced_dev(int mode, const double* a33, const double* b33, const double* nlm, int64_t ldn, int64_t N, int64_t ld,
                        double iota, double zeta, double* Mrr, double* Mri, double* Mir, double* Mii, void* stream) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!a33 || !Mrr || !Mri || !Mir || !Mii || (mode == 0 && !b33) || (mode == 2 && (!nlm || ldn < N))) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_mexport_reduced(mode, g.L, a33, b33, reinterpret_cast<const double2*>(nlm), ldn, N, ld, iota, zeta, Mrr, Mri, Mir, Mii,
                                  (cudaStream_t)stream));
    return SFB_OK;
}
static int mreduced_host(int mode, const double* a33, const double* b33, const double* nlm, int64_t ldn, int64_t N, double iota,
                         double zeta, double* Mrr, double* Mri, double* Mir, double* Mii) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!a33 || !Mrr || !Mri || !Mir || !Mii || (mode == 0 && !b33) || (mode == 2 && !nlm)) return fail(SFB_EINVAL, "null array");
    DevTmp da, db, dn, dm;
    int rc;
    CK(da.alloc((size_t)N * 72));
    CK(cudaMemcpy(da.p, a33, (size_t)N * 72, cudaMemcpyHostToDevice));
    if (mode == 0) { CK(db.alloc((size_t)N * 72)); CK(cudaMemcpy(db.p, b33, (size_t)N * 72, cudaMemcpyHostToDevice)); }
    if (mode == 2 && (rc = stage_rows(dn, nlm, N, ldn, 15))) return rc;
    const int r = sfb_rnlm_len();
    const size_t one = (size_t)N * r * r;
    CK(dm.alloc(4 * one * 8));
    double* o = dm.as<double>();
    rc = mreduced_dev(mode, da.as<double>(), db.as<double>(), dn.as<double>(), N, N, N, iota, zeta, o, o + one, o + 2 * one, o + 3 * one, nullptr);
    if (rc) return rc;
    double* dst[4] = {Mrr, Mri, Mir, Mii};
    for (int q = 0; q < 4; ++q) CK(cudaMemcpy(dst[q], o + q * one, one * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_M_LROT_reduced_arr(const double* eps, const double* omg, int64_t N, double iota, double zeta,
                           double* Mrr, double* Mri, double* Mir, double* Mii) {
    return mreduced_host(0, eps, omg, nullptr, 0, N, iota, zeta, Mrr, Mri, Mir, Mii);
}
int sfb_M_LROT_reduced_arr_dev(const double* eps, const double* omg, int64_t N, int64_t ld, double iota, double zeta,
                               double* Mrr, double* Mri, double* Mir, double* Mii, void* stream) {
    return mreduced_dev(0, eps, omg, nullptr, 0, N, ld, iota, zeta, Mrr, Mri, Mir, Mii, stream);
}
int sfb_M_DDRX_reduced_arr(const double* nlm, int64_t ldn, const double* tau, int64_t N, int src_only,
                           double* Mrr, double* Mri, double* Mir, double* Mii) {
    return mreduced_host(src_only ? 1 : 2, tau, nullptr, nlm, ldn, N, 0, 0, Mrr, Mri, Mir, Mii);
}
int sfb_M_DDRX_reduced_arr_dev(const double* nlm, int64_t ldn, const double* tau, int64_t N, int64_t ld, int src_only,
                               double* Mrr, double* Mri, double* Mir, double* Mii, void* stream) {
    return mreduced_dev(src_only ? 1 : 2, tau, nullptr, nlm, ldn, N, ld, 0, 0, Mrr, Mri, Mir, Mii, stream);
}
int sfb_reduce_M_arr_dev(const double* M, int is_complex, int64_t N, int64_t ld, double* Mrr, double* Mri, double* Mir, double* Mii,
                         void* stream) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!M || !Mrr || !Mri || !Mir || !Mii) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_reduce_dense(M, is_complex, g.L, N, ld, Mrr, Mri, Mir, Mii, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_reduce_M_arr(const double* M, int is_complex, int64_t N, double* Mrr, double* Mri, double* Mir, double* Mii) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!M || !Mrr || !Mri || !Mir || !Mii) return fail(SFB_EINVAL, "null array");
    DevTmp di, dm;
    const size_t ibytes = (size_t)N * g.n * g.n * (is_complex ? 16 : 8);
    CK(di.alloc(ibytes));
    CK(cudaMemcpy(di.p, M, ibytes, cudaMemcpyHostToDevice));
    const int r = sfb_rnlm_len();
    const size_t one = (size_t)N * r * r;
    CK(dm.alloc(4 * one * 8));
    double* o = dm.as<double>();
    int rc = sfb_reduce_M_arr_dev(di.as<double>(), is_complex, N, N, o, o + one, o + 2 * one, o + 3 * one, nullptr);
    if (rc) return rc;
    double* dst[4] = {Mrr, Mri, Mir, Mii};
    for (int q = 0; q < 4; ++q) CK(cudaMemcpy(dst[q], o + q * one, one * 8, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_M_REG_arr_dev(const double* eps, int64_t N, int64_t ld, double* M, void* stream) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0 || ld < N) return fail(SFB_EINVAL, "need 0 <= N <= ld");
    if (N == 0) return SFB_OK;
    if (!eps || !M) return fail(SFB_EINVAL, "null array");
    CK(sfb_launch_mreg(g.L, g.reg, eps, N, ld, M, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_M_REG_arr(const double* eps, int64_t N, double* M) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!eps || !M) return fail(SFB_EINVAL, "null array");
    DevTmp de, dm;
    CK(de.alloc((size_t)N * 72));
    CK(cudaMemcpy(de.p, eps, (size_t)N * 72, cudaMemcpyHostToDevice));
    const size_t mbytes = (size_t)N * g.n * g.n * 8;
    CK(dm.alloc(mbytes));
    int rc = sfb_M_REG_arr_dev(de.as<double>(), N, N, dm.as<double>(), nullptr);
    if (rc) return rc;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(M, dm.p, mbytes, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_M_CDRX(double* M) {      // constant operator diag(-l(l+1)), src/dynamics.f90:474-492 (no per-node input)
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (!M) return fail(SFB_EINVAL, "null array");
    memset(M, 0, sizeof(double) * g.n * g.n);
    int j = 0;
    for (int l = 0; l <= g.L; l += 2)
        for (int m = -l; m <= l; ++m, ++j) M[(size_t)j * g.n + j] = -(double)(l * (l + 1));
    return SFB_OK;
}

// ---------------------------------------------------------------------------------------------
// apply_bounds, reduced form
// ---------------------------------------------------------------------------------------------
int sfb_rnlm_len(void) { return g.L ? (g.L + 2) * (g.L + 2) / 4 : 0; }          // src/reducedform.f90:26

int sfb_apply_bounds_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out, void* stream) {
    int rc = basic_check(nlm_in, N, ld_in);
    if (rc) return rc;
    if (ld_out < N || (N && !nlm_out)) return fail(SFB_EINVAL, "bad output");
    CK(sfb_launch_bounds(reinterpret_cast<const double2*>(nlm_in), reinterpret_cast<double2*>(nlm_out), N, ld_in, ld_out, g.n, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_apply_bounds_arr(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld) {
    int rc = basic_check(nlm_in, N, ld);
    if (rc || N == 0) return rc;
    if (!nlm_out) return fail(SFB_EINVAL, "null output");
    DevTmp in;                      // only the l <= 4 rows change: stage 15 rows, rescale in place, copy them back
    if ((rc = stage_rows(in, nlm_in, N, ld, 15))) return rc;
    CK(sfb_launch_bounds(in.as<double2>(), in.as<double2>(), N, N, N, 15, nullptr));
    if (nlm_out != nlm_in) {
        for (int j = 15; j < g.n; ++j) memcpy(nlm_out + 2 * (size_t)j * ld, nlm_in + 2 * (size_t)j * ld, (size_t)N * 16);
    }
    CK(cudaMemcpy2D(nlm_out, (size_t)ld * 16, in.p, (size_t)N * 16, (size_t)N * 16, 15, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_apply_bounds_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out, void* stream) {
    int rc = basic_check(rnlm_in, N, ld_in);
    if (rc) return rc;
    if (ld_out < N || (N && !rnlm_out)) return fail(SFB_EINVAL, "bad output");
    CK(sfb_launch_bounds_rnlm(reinterpret_cast<const double2*>(rnlm_in), reinterpret_cast<double2*>(rnlm_out), N, ld_in, ld_out,
                              sfb_rnlm_len(), (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_nlm_to_rnlm_arr_dev(const double* nlm, double* rnlm, int64_t N, int64_t ld_nlm, int64_t ld_rnlm, void* stream) {
    int rc = basic_check(nlm, N, ld_nlm);
    if (rc) return rc;
    if (ld_rnlm < N || (N && !rnlm)) return fail(SFB_EINVAL, "bad output");
    CK(sfb_launch_reduced(1, reinterpret_cast<const double2*>(nlm), reinterpret_cast<double2*>(rnlm), N, ld_nlm, ld_rnlm, g.L, (cudaStream_t)stream));
    return SFB_OK;
}
int sfb_rnlm_to_nlm_arr_dev(const double* rnlm, double* nlm, int64_t N, int64_t ld_rnlm, int64_t ld_nlm, void* stream) {
    int rc = basic_check(rnlm, N, ld_rnlm);
    if (rc) return rc;
    if (ld_nlm < N || (N && !nlm)) return fail(SFB_EINVAL, "bad output");
    CK(sfb_launch_reduced(0, reinterpret_cast<const double2*>(rnlm), reinterpret_cast<double2*>(nlm), N, ld_rnlm, ld_nlm, g.L, (cudaStream_t)stream));
    return SFB_OK;
}
static int reduced_host(int to_reduced, const double* src, double* dst, int64_t N) {
    if (!g.L) return fail(SFB_ENOINIT, "sfb_init not called");
    if (N < 0) return fail(SFB_EINVAL, "N < 0");
    if (N == 0) return SFB_OK;
    if (!src || !dst) return fail(SFB_EINVAL, "null array");
    const size_t nr = (size_t)sfb_rnlm_len(), ns = to_reduced ? (size_t)g.n : nr, nd = to_reduced ? nr : (size_t)g.n;
    DevTmp a, b;
    CK(a.alloc(ns * N * 16));
    CK(b.alloc(nd * N * 16));
    CK(cudaMemcpy(a.p, src, ns * N * 16, cudaMemcpyHostToDevice));
    CK(sfb_launch_reduced(to_reduced, a.as<double2>(), b.as<double2>(), N, N, N, g.L, nullptr));
    CK(cudaMemcpy(dst, b.p, nd * N * 16, cudaMemcpyDeviceToHost));
    return SFB_OK;
}
int sfb_nlm_to_rnlm_arr(const double* nlm, double* rnlm, int64_t N) { return reduced_host(1, nlm, rnlm, N); }
int sfb_rnlm_to_nlm_arr(const double* rnlm, double* nlm, int64_t N) { return reduced_host(0, rnlm, nlm, N); }

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
int sfb_host_register(void* p, int64_t bytes) {
    if (!p || bytes <= 0) return fail(SFB_EINVAL, "sfb_host_register: null pointer or empty range");
    CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable));
    return SFB_OK;
}
int sfb_host_unregister(void* p) {
    if (!p) return fail(SFB_EINVAL, "sfb_host_unregister: null pointer");
    CK(cudaHostUnregister(p));
    return SFB_OK;
}
int sfb_sync(void) { CK(cudaDeviceSynchronize()); return SFB_OK; }
int sfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int sfb_set_device(int dev) { CK(cudaSetDevice(dev)); return SFB_OK; }

}  // extern "C"
