// Fused fabric-evolution step kernel, FOUR lanes per node (sm_100a), one translation unit per (L, term set).
//
//   d nlm/dt = (M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + M_REG) nlm      Euler / classical RK4
//   (reference: src/dynamics.f90:52-97, 251-298, 402-422, 474-518, 108; RK4 per BASELINE config 2)
//
// Work decomposition (see codegen/emit_step.py, emit4): a node is served by the lanes
//   (sign set A: rows m>=0 | B: rows m<=0)  x  (component: re | im)
// which all execute ONE generated straight-line instruction stream; a warp holds 8 nodes, a CTA
// SFB_TN nodes = SFB_TN/8 warps kept in lock step (one instruction-cache window for the whole SM:
// profiles/r01_notes.md).  The tile's state lives in shared memory for all RK stages (1-D TMA bulk
// loads, one per coefficient row); every state value is read from shared memory once per stage and
// kept in registers across the canonical-m sweep; table entries are immediates.
//
// Included by generated .cu files that define:
//   SFB_L, SFB_DDRX (0/1), SFB_TN, SFB_MINB, SFB_NAME (launcher symbol), SFB_APPLY_INC
#pragma once
#include <cstring>
#include <mutex>
#include "sfb_common.cuh"
#include "sfb_moments.cuh"

namespace {

constexpr int kL = SFB_L;
constexpr int kNCoef = (kL + 1) * (kL + 2) / 2;
constexpr int kNRow = 2 * (kL / 2 + 1) * (kL / 2 + 1);   // physical rows: plane pair (m>=0 | m<0) per (l,|m|) slot
constexpr int kTN = SFB_TN;
constexpr int kThreads = 4 * kTN;
constexpr int kNF = SFB_DDRX ? 23 : 8;                    // forcing entries per lane set
#define SFB_NO_DSCAL 1                                   // <D> ingredients are recomputed per stage, not kept in shared memory
constexpr int kNSc = 4;                                   // per-node scalars (SC_C0, SC_LAM, SC_RM, SC_G0)
static_assert(kTN % 8 == 0, "tile must be a multiple of 8 nodes (one warp)");

__constant__ SfbRegConst c_reg;

__host__ __device__ constexpr int pslot(int l, int a) { return (l / 2) * (l / 2) + a; }
__host__ __device__ constexpr int hrow(int l) { return l * (l + 1) / 2; }

enum { SC_C0 = 0, SC_LAM = 1, SC_RM = 2, SC_G0 = 3, SC_TAUV = 4, SC_TSQV = 10, SC_NORM = 16 };

struct Ctx {
    const double *yz, *yp, *yn;       // stage input planes for this lane's component (zero / pos / neg canonical m)
    const double2* ktab;              // global table of operator entries (gtab mode)
    const double2* fz;                // forcing block of this lane's sign set
    double *oz, *op;                  // next-stage buffer (rows owned by this lane)
    double *az, *ap;                  // RK accumulator buffer
    double* gout;                     // global output: component of node, row stride 2*ld doubles
    const double* gin;                // global input (n0 re-read in RK stages 2,3)
    long long ld_out, sld, ld_in, sld_in;
    double c0, lam, rm;               // diagonal: c0 + lam*(-l(l+1)) + rm*regdiag_l
    double as, bs;                    // ynext = n0 + as*k ; acc += bs*k
    double sigma;                     // -1 (re lane) / +1 (im lane): sign of the partner's contribution
    bool first, last, isA, valid, ld_n0, ld_acc;
};

// RK4 formulation.  Kernels without DDRX are linear in nlm with stage-independent M, for which classical
// RK4 equals the 4th-order Taylor polynomial; it is evaluated in Horner form
//     y1 = n0 + dt/4 M n0 ; y2 = n0 + dt/3 M y1 ; y3 = n0 + dt/2 M y2 ; out = n0 + dt M y3
// (2 state buffers, no accumulator; differs from the k1..k4 form only by roundings, ~1e-16).
// DDRX kernels (<D> depends on the stage state) use the classical k1..k4 form with an accumulator buffer.
#define SFB_HORNER (!SFB_DDRX)

// loads issued at the START of a canonical-m block so that their latency overlaps the block's arithmetic
template <int l, int mu>
__device__ __forceinline__ double n0_load(const Ctx& c) {
    double v = 0.0;
    if (c.ld_n0 && (mu != 0 || c.isA)) v = c.gin[2 * ((long long)hrow(l) * c.ld_in + (long long)mu * c.sld_in)];
    return v;
}
template <int l, int mu>
__device__ __forceinline__ double acc_load(const Ctx& c) {
#if SFB_HORNER
    return 0.0;
#else
    double v = 0.0;
    if (c.ld_acc && (mu != 0 || c.isA)) v = (mu == 0 ? c.az : c.ap)[4 * pslot(l, mu) * kTN];   // set B never owns m = 0 rows
    return v;
#endif
}
// finalize one row (branch-free: selects + predicated stores): exchange the cross-component partial sum
// with the partner lane, add the diagonal terms, apply the stage update for this lane's component
template <int l, int mu>
__device__ __forceinline__ void row_out(const Ctx& c, double mine, double theirs, double z, double n0, double acc) {
    const double recv = __shfl_xor_sync(0xffffffffu, theirs, 1);
    double k = fma(c.sigma, recv, mine);
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    k = fma(d, z, k);
    constexpr int off = 4 * pslot(l, mu) * kTN;      // doubles
    const bool own = (mu != 0) || c.isA;             // m = 0 rows are computed by both sign sets; A owns them
    const double n0v = c.first ? z : n0;
    const long long goff = 2 * ((long long)hrow(l) * c.ld_out + (long long)mu * c.sld);
#if SFB_HORNER
    const double y = fma(c.as, k, n0v);
    if (own && !c.last) (mu == 0 ? c.oz : c.op)[off] = y;
    if (own && c.last && c.valid) c.gout[goff] = y;
#else
    const double A = fma(c.bs, k, c.first ? z : acc);
    const double y = fma(c.as, k, n0v);
    if (own && !c.last) { (mu == 0 ? c.oz : c.op)[off] = y; (mu == 0 ? c.az : c.ap)[off] = A; }
    if (own && c.last && c.valid) c.gout[goff] = A;
#endif
}
#define SFB_ROW_OUT4(l, mu, m, t, z, q, r) row_out<l, mu>(c, m, t, z, q, r)
#define SFB_N0_LOAD(l, mu) n0_load<l, mu>(c)
#define SFB_ACC_LOAD(l, mu) acc_load<l, mu>(c)
#define SFB_LOCKSTEP(lo, hi) __syncthreads()

__device__ __forceinline__ void apply_all(const Ctx& c) {
    const double* __restrict__ yz = c.yz;
    const double* __restrict__ yp = c.yp;
    const double* __restrict__ yn = c.yn;
    const double2* __restrict__ fz = c.fz;
#ifdef SFB_GTAB
    const double2* __restrict__ ktab = c.ktab;     // table entries: uniform-address 128-bit loads through L1
#endif
#include SFB_APPLY_INC
}

#include "sfb_step_common.cuh"

// the whole step of one tile of kTN nodes starting at node0 (kThreads = 4*kTN threads).  Also the fallback of the reduced
// two-lane kernel (sfb_step_kernel4r.cuh) for tiles whose states lack the real-ODF symmetry.
#ifdef SFB_REDUCED
#define SFB_TILE_FN __device__ __noinline__
#else
#define SFB_TILE_FN __device__ __forceinline__
#endif
SFB_TILE_FN void full_tile(const SfbStepParams& P, const long long node0, unsigned char* smem_raw) {
    const int nbuf = P.nstage == 1 ? 1 : (SFB_HORNER ? 2 : 3);
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRow * kTN;
    double* scal = reinterpret_cast<double*>(forc + 2 * kNF * kTN);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(scal + kNSc * kTN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sb = lane >> 4;                          // 0: sign set A (m>=0), 1: set B (m<=0)
    const int comp = lane & 1;                         // 0: real part, 1: imaginary part
    const int nl = warp * 8 + ((lane & 15) >> 1);      // node within tile
    const int nvalid = (int)min((long long)kTN, P.N - node0);

    const uint32_t mb = smem_u32(mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t bytes = (uint32_t)nvalid * 16u;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes * (uint32_t)kNCoef) : "memory");
        for (int j = lane; j < kNCoef; j += 32) {
            int l = 0;
            while ((l + 2) * (l + 3) / 2 - (l + 2) <= j) l += 2;
            const int m = j - hrow(l);
            const int prow = 2 * pslot(l, m < 0 ? -m : m) + (m < 0 ? 1 : 0);
            const uint32_t dst = smem_u32(bufs + (size_t)prow * kTN);
            const double2* src = P.nlm_in + (long long)j * P.ld_in + node0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    }
    prep_tile(P, node0, nvalid, tid, forc, scal);
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }
    __syncthreads();

    Ctx c;
    c.isA = (sb == 0);
#ifdef SFB_GTAB
    c.ktab = P.ktab;
#else
    c.ktab = nullptr;
#endif
    c.valid = nl < nvalid;
    c.sigma = comp ? 1.0 : -1.0;
    c.fz = forc + (size_t)sb * kNF * kTN + nl;
    c.lam = scal[SC_LAM * kTN + nl];
    c.rm = scal[SC_RM * kTN + nl];
    c.c0 = 0.0;
    c.ld_out = P.ld_out; c.sld = sb ? -P.ld_out : P.ld_out;
    c.ld_in = P.ld_in;   c.sld_in = sb ? -P.ld_in : P.ld_in;
    c.gout = reinterpret_cast<double*>(P.nlm_out + node0 + nl) + comp;
    c.gin = reinterpret_cast<const double*>(P.nlm_in + node0 + nl) + comp;
    double* dbuf = reinterpret_cast<double*>(bufs) + 2 * nl + comp;           // this lane's component column
    constexpr size_t kBufD = (size_t)2 * kNRow * kTN;                          // doubles per buffer
    constexpr int kPlane = 2 * kTN;                                            // doubles per physical row
    double* acc = dbuf + (size_t)(nbuf - 1) * kBufD;
    c.az = acc; c.ap = acc + sb * kPlane;

    for (int s = 0; s < P.nstage; ++s) {
        // RK4 (3 buffers, n0 re-read from global): inputs 0,1,0,1 ; outputs 1,0,1 ; accumulator 2
        const int ib = s & 1, ob = (s + 1) & 1;
        const double* yin = dbuf + (size_t)ib * kBufD;
        double* yout = dbuf + (size_t)ob * kBufD;
        c.yz = yin; c.yp = yin + sb * kPlane; c.yn = yin + (1 - sb) * kPlane;
        c.oz = yout; c.op = yout + sb * kPlane;
        c.first = (s == 0);
        c.last = (s == P.nstage - 1);
        c.ld_n0 = !c.first && c.valid && (SFB_HORNER || !c.last);
        c.ld_acc = !c.first;
#if SFB_HORNER
        c.as = (P.nstage == 1) ? P.dt : P.dt / (double)(4 - s);
        c.bs = 0.0;
#else
        if (P.nstage == 1) { c.as = 0.0; c.bs = P.dt; }
        else {
            c.as = (s == 2) ? P.dt : 0.5 * P.dt;
            c.bs = (s == 0 || s == 3) ? P.dt / 6 : P.dt / 3;
        }
#endif
#if SFB_DDRX
        if (tid < nvalid) {   // <D>(current stage state), one thread per node
            const double2* y = bufs + (size_t)ib * kNRow * kTN + tid;
            double2 n2[3], n4[5];
#pragma unroll
            for (int m = 0; m < 3; ++m) n2[m] = y[2 * pslot(2, m) * kTN];
#pragma unroll
            for (int m = 0; m < 5; ++m) n4[m] = (kL >= 4) ? y[2 * pslot(4, m) * kTN] : make_double2(0.0, 0.0);
            const double davg = ddrx_davg(global_src(P, node0 + tid), y[0], n2, n4);
            scal[SC_C0 * kTN + tid] = -(scal[SC_G0 * kTN + tid] * davg);
        }
        __syncthreads();
        c.c0 = scal[SC_C0 * kTN + nl];
#endif
        apply_all(c);
        if (!c.last) __syncthreads();
    }
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");   // full_tile may run again in this CTA
}

#ifndef SFB_REDUCED
__global__ void __launch_bounds__(kThreads, SFB_MINB) step_kernel(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    full_tile(P, (long long)blockIdx.x * kTN, smem_raw);
}
#endif

}  // namespace

#ifdef SFB_REDUCED
#include "sfb_step_kernel4r.cuh"
#else

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    const size_t fixed = (size_t)2 * kNF * kTN * 16 + (size_t)kNSc * kTN * 8 + 16;
    const size_t per_buf = (size_t)kNRow * kTN * 16;
    const size_t lim = 227 * 1024;
    const int nbuf_rk = SFB_HORNER ? 2 : 3;
    const size_t smem_max = (nbuf_rk * per_buf + fixed <= lim) ? nbuf_rk * per_buf + fixed : per_buf + fixed;
    cudaError_t e;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::unique_lock<std::mutex> attr_lk(attr_mu);
    if (!attr_done[dev]) {
        e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    attr_lk.unlock();
    if (Pin.rio) return cudaErrorNotSupported;      // reduced-form arrays: reduced kernels only
    SfbStepParams P = Pin;
#ifdef SFB_GTAB
    { void* tp = nullptr; e = cudaGetSymbolAddress(&tp, sfb_gtab); if (e != cudaSuccess) return e; P.ktab = reinterpret_cast<const double2*>(tp); }
#endif
    P.n0_global = 1;
    const int nbuf = P.nstage == 1 ? 1 : nbuf_rk;
    const size_t smem = nbuf * per_buf + fixed;
    if (smem > lim) return cudaErrorInvalidConfiguration;
    {
        // uploaded once per device under a lock, and WAITED for: a later launch on another stream carries no dependency on this copy
        static std::mutex reg_mu;
        static SfbRegConst last[64];
        static bool have[64] = {false};
        std::lock_guard<std::mutex> reg_lk(reg_mu);
        if (!have[dev] || memcmp(&last[dev], &reg, sizeof(SfbRegConst)) != 0) {
            e = cudaMemcpyToSymbolAsync(c_reg, &reg, sizeof(SfbRegConst), 0, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            last[dev] = reg;
            have[dev] = true;
        }
    }
    if (P.N <= 0) return cudaSuccess;
    const long long ntile = (P.N + kTN - 1) / kTN;
    step_kernel<<<(unsigned)ntile, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}
#endif  // SFB_REDUCED
