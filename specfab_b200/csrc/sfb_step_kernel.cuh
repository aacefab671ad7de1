// Fused fabric-evolution step kernel (sm_100a), one translation unit per (L, term set).
//
// Computes, for a tile of SFB_TN nodes per CTA,  nlm <- step(nlm)  with
//     d nlm/dt = (M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + M_REG) nlm
// (reference: src/dynamics.f90:52-97, 251-298, 402-422, 474-518; Euler update src/dynamics.f90:108;
//  classical RK4 per BASELINE config 2) without ever forming M:
//   * the tile's state is staged once in shared memory by 1-D TMA bulk copies (one per coefficient
//     row, rows are node-contiguous) and stays there for all RK stages;
//   * every node is served by 2*SFB_R lanes: lane set A (rows m>=0) / B (rows m<=0) of SFB_R warp
//     roles; A and B run the same generated instruction stream (mirror symmetry of the Gaunt
//     tables, see codegen/operators.py), roles split the canonical m's;
//   * table entries are immediates of the generated straight-line code (codegen/emit_step.py).
//
// Included by generated .cu files that define:
//   SFB_L, SFB_DDRX (0/1), SFB_R, SFB_TN, SFB_MINB, SFB_NAME (launcher symbol), SFB_APPLY_INC
#pragma once
#include <cstring>
#include <mutex>
#include "sfb_common.cuh"
#include "sfb_moments.cuh"

namespace {

constexpr int kL = SFB_L;
constexpr int kNCoef = (kL + 1) * (kL + 2) / 2;
constexpr int kNRow = 2 * (kL / 2 + 1) * (kL / 2 + 1);   // physical rows: 2 planes (m>=0 | m<0) per (l,|m|) slot
constexpr int kTN = SFB_TN;
constexpr int kR = SFB_R;
constexpr int kG = kTN / 16;
constexpr int kThreads = 32 * kG * kR;
constexpr int kNF = SFB_DDRX ? 23 : 8;                    // forcing entries per lane set
#define SFB_NO_DSCAL 1                                   // <D> ingredients are recomputed per stage, not kept in shared memory
constexpr int kNSc = 4;                                   // per-node scalars (SC_C0, SC_LAM, SC_RM, SC_G0)
static_assert(kTN % 16 == 0, "tile must be a multiple of 16 nodes");
// SFB_CW > 1 (reduced one-lane kernel, one role): a CTA holds SFB_CW independent one-warp tiles -- warps that start
// together and run the same instruction stream share its instruction-cache misses.  Every tile-level barrier is then a
// warp barrier and every tile has its own slice of the dynamic shared memory.
#ifndef SFB_CW
#define SFB_CW 1
#endif
constexpr int kCW = SFB_CW;
static_assert(kCW == 1 || kThreads == 32, "several tiles per CTA: one-warp tiles only");
#define SFB_TILE_SYNC() do { if (kCW > 1) __syncwarp(); else __syncthreads(); } while (0)

__constant__ SfbRegConst c_reg;

__host__ __device__ constexpr int pslot(int l, int a) { return (l / 2) * (l / 2) + a; }
__host__ __device__ constexpr int hrow(int l) { return l * (l + 1) / 2; }

// scalar slots
enum { SC_C0 = 0, SC_LAM = 1, SC_RM = 2, SC_G0 = 3, SC_TAUV = 4, SC_TSQV = 10, SC_NORM = 16 };

struct Ctx {
    const double2 *yz, *yp, *yn;      // stage input planes (zero / positive / negative canonical m)
    const double2* ktab;              // global table of operator entries (gtab / loop mode)
    double2* ring;                    // loop mode: this warp's private table ring in shared memory
    const double2* fz;                // forcing block of this lane set
    double2 *oz, *op;                 // next-stage buffer (rows owned by this lane: zero / positive plane)
    double2 *az, *ap;                 // RK accumulator buffer
    double2* gout;                    // global output, already offset by node
    const double2* gin;               // global input (n0 re-read when the 4th smem buffer does not fit)
    long long ld_out, sld;            // row stride, signed row stride (+ld for set A, -ld for set B)
    long long ld_in, sld_in;
    double c0, lam, rm;               // diagonal: c0 + lam*(-l(l+1)) + rm*regdiag_l
    double as, bs;                    // stage coefficients: ynext = n0 + as*k ; acc += bs*k
    bool first, last, isA, valid, ld_n0, ld_acc;
};

// stage buffers of a multi-stage step: ping-pong pair (+ accumulator for the classical form), or with SFB_INPLACE one
// buffer updated in place through a register delay queue (codegen/emit_step.py::emit, single role only)
#ifdef SFB_INPLACE
constexpr bool kInPlace = true;
#else
constexpr bool kInPlace = false;
#endif
constexpr int kNBufRK = (kInPlace ? 1 : 2) + (SFB_DDRX ? 1 : 0);

// RK4 formulation: see sfb_step_kernel4.cuh (Horner form of the Taylor polynomial for the linear,
// DDRX-free kernels; classical k1..k4 with an accumulator buffer when DDRX makes the RHS state dependent).
#define SFB_HORNER (!SFB_DDRX)

// loads issued at the START of a canonical-m block so that their latency overlaps the block's arithmetic
template <int l, int mu>
__device__ __forceinline__ double2 n0_load(const Ctx& c) {
    double2 v = make_double2(0.0, 0.0);
    if (c.ld_n0 && (mu != 0 || c.isA)) v = c.gin[(long long)hrow(l) * c.ld_in + (long long)mu * c.sld_in];
    return v;
}
template <int l, int mu>
__device__ __forceinline__ double2 acc_load(const Ctx& c) {
    double2 v = make_double2(0.0, 0.0);
#if !SFB_HORNER
    if (c.ld_acc && (mu != 0 || c.isA)) v = (mu == 0 ? c.az : c.ap)[2 * pslot(l, mu) * kTN];   // set B never owns m = 0 rows
#endif
    return v;
}
// finalize one row, branch-free (selects + predicated stores); returns the stage output y.  STORE: write y to the
// next-stage buffer now; otherwise the caller commits it later (in-place scheme)
template <int l, int mu, bool STORE>
__device__ __forceinline__ double2 row_out(const Ctx& c, double kr, double ki, double zr, double zi, double2 n0, double2 acc) {
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    kr = fma(d, zr, kr);
    ki = fma(d, zi, ki);
    constexpr int off = 2 * pslot(l, mu) * kTN;
    const bool own = (mu != 0) || c.isA;             // m = 0 rows are computed by both lane sets; set A owns them
    const double n0r = c.first ? zr : n0.x, n0i = c.first ? zi : n0.y;
    const long long goff = (long long)hrow(l) * c.ld_out + (long long)mu * c.sld;
#if SFB_HORNER
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (STORE && own && !c.last) (mu == 0 ? c.oz : c.op)[off] = y;
    if (own && c.last && c.valid) c.gout[goff] = y;
#else
    const double2 A = make_double2(fma(c.bs, kr, c.first ? zr : acc.x), fma(c.bs, ki, c.first ? zi : acc.y));
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (own && !c.last) { if (STORE) (mu == 0 ? c.oz : c.op)[off] = y; (mu == 0 ? c.az : c.ap)[off] = A; }
    if (own && c.last && c.valid) c.gout[goff] = A;
#endif
    return y;
}
template <int l, int mu>
__device__ __forceinline__ void row_commit(const Ctx& c, double2 y) {
    const bool own = (mu != 0) || c.isA;
    if (own && !c.last) (mu == 0 ? c.oz : c.op)[2 * pslot(l, mu) * kTN] = y;
}
#define SFB_ROW_PRE(l, mu, q, r) const double2 q = n0_load<l, mu>(c), r = acc_load<l, mu>(c)
#define SFB_ROW_OUT(l, mu, ar, ai, zr, zi, q, r) row_out<l, mu, true>(c, ar, ai, zr, zi, q, r)
#define SFB_ROW_OUTQ(l, mu, ar, ai, zr, zi, q, r, o) o = row_out<l, mu, false>(c, ar, ai, zr, zi, q, r)
#define SFB_ROW_COMMIT(l, mu, o) row_commit<l, mu>(c, o)
#define SFB_COMMIT_FENCE() __syncwarp()      // the partner lane set has read the rows that are overwritten next (one-warp CTAs)
// keeps the warps of a CTA within one instruction-cache window of the generated straight-line code
#define SFB_LOCKSTEP(lo, hi) __syncthreads()

#ifdef SFB_LOOP
#include "sfb_step_loop.cuh"
__device__ __forceinline__ void apply_role(const Ctx& c, int role) { loopk::apply_loop(c, role, kR, c.ring, (int)(threadIdx.x & 31)); }
constexpr size_t kRingBytes = (size_t)(32 * (SFB_TN / 16) * SFB_R / 32) * 2 * (SFB_LT_PER_CHUNK / 2) * 16;
#else
constexpr size_t kRingBytes = 0;
__device__ __forceinline__ void apply_role(const Ctx& c, int role) {
    const double2* __restrict__ yz = c.yz;
    const double2* __restrict__ yp = c.yp;
    const double2* __restrict__ yn = c.yn;
    const double2* __restrict__ fz = c.fz;
#ifdef SFB_GTAB
    const double2* __restrict__ ktab = c.ktab;     // table entries: uniform-address 128-bit loads through L1
#endif
#include SFB_APPLY_INC
}
#endif

#include "sfb_step_common.cuh"

// the whole step of one tile of kTN nodes starting at node0 (the CTA's kThreads threads cooperate).  Also the
// fallback of the reduced kernel (sfb_step_kernel_r.cuh) for tiles whose states lack the real-ODF symmetry.
#ifdef SFB_REDUCED
#define SFB_TILE_FN __device__ __noinline__      // rarely taken: keep it out of the reduced kernel's register allocation
#else
#define SFB_TILE_FN __device__ __forceinline__
#endif
SFB_TILE_FN void full_tile(const SfbStepParams& P, const long long node0, unsigned char* smem_raw) {
    const int nbuf = P.nstage == 1 ? 1 : kNBufRK;
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRow * kTN;
    double* scal = reinterpret_cast<double*>(forc + 2 * kNF * kTN);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(scal + kNSc * kTN);
    double2* rings = reinterpret_cast<double2*>(mbar + 2);          // loop mode: [warps][2][pairs per item]

    const int tid = threadIdx.x % kThreads, warp = tid >> 5, lane = tid & 31;
    const int group = warp / kR, role = warp % kR;
    const int sb = lane >> 4;                       // 0: lane set A (m>=0), 1: lane set B (m<=0)
    const int nl = group * 16 + (lane & 15);        // node within tile
    const int nvalid = (int)min((long long)kTN, P.N - node0);

    // ---- stage the tile: one bulk copy per coefficient row into buffer 0
    const uint32_t mb = smem_u32(mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    SFB_TILE_SYNC();
    if (warp == 0) {
        const uint32_t bytes = (uint32_t)nvalid * 16u;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes * (uint32_t)kNCoef) : "memory");
        for (int j = lane; j < kNCoef; j += 32) {
            // (l,m) of global row j:  j = l(l+1)/2 + m, even l
            int l = 0;
            while ((l + 2) * (l + 3) / 2 - (l + 2) <= j) l += 2;   // first row of degree l+2 is hrow(l+2)-(l+2)
            const int m = j - hrow(l);
            const int prow = 2 * pslot(l, m < 0 ? -m : m) + (m < 0 ? 1 : 0);
            const uint32_t dst = smem_u32(bufs + (size_t)prow * kTN);
            const double2* src = P.nlm_in + (long long)j * P.ld_in + node0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    }
#ifdef SFB_LOOP
    // the loop kernel reads columns unpredicated: the never-written twin rows of the m = 0 slots must be finite
    for (int b = 0; b < nbuf; ++b)
        for (int h = 0; h <= kL / 2; ++h)
            for (int t = tid; t < kTN; t += kThreads) bufs[((size_t)b * kNRow + 2 * h * h + 1) * kTN + t] = make_double2(0.0, 0.0);
#endif
    // ---- meanwhile: per-node forcing
    prep_tile(P, node0, nvalid, tid, forc, scal);
    {   // wait for the tile
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }
    SFB_TILE_SYNC();

    Ctx c;
    c.isA = (sb == 0);
#ifdef SFB_LOOP
    c.ring = rings + (size_t)warp * 2 * (SFB_LT_PER_CHUNK / 2);
#else
    c.ring = nullptr;
#endif
#if defined(SFB_GTAB) || defined(SFB_LOOP)
    c.ktab = P.ktab;
#else
    c.ktab = nullptr;
#endif
    c.valid = nl < nvalid;
    c.fz = forc + (size_t)sb * kNF * kTN + nl;
    c.lam = scal[SC_LAM * kTN + nl];
    c.rm = scal[SC_RM * kTN + nl];
    c.c0 = 0.0;
    c.ld_out = P.ld_out;
    c.sld = sb ? -P.ld_out : P.ld_out;
    c.gout = P.nlm_out + node0 + nl;
    c.gin = P.nlm_in + node0 + nl;
    c.ld_in = P.ld_in;
    c.sld_in = sb ? -P.ld_in : P.ld_in;
    double2* acc = bufs + (size_t)(nbuf - 1) * kNRow * kTN + nl;
    c.az = acc; c.ap = acc + sb * kTN;

    for (int s = 0; s < P.nstage; ++s) {
        // n0 is re-read from global (L2) in the later RK stages: inputs 0,1,0,1 ; outputs 1,0,1 ; accumulator 2
        const int ib = kInPlace ? 0 : (s & 1), ob = kInPlace ? 0 : ((s + 1) & 1);
        const double2* yin = bufs + (size_t)ib * kNRow * kTN + nl;
        double2* yout = bufs + (size_t)ob * kNRow * kTN + nl;
        c.yz = yin; c.yp = yin + sb * kTN; c.yn = yin + (1 - sb) * kTN;
        c.oz = yout; c.op = yout + sb * kTN;
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
        SFB_TILE_SYNC();
        c.c0 = scal[SC_C0 * kTN + nl];
#endif
        apply_role(c, role);
        if (!c.last) SFB_TILE_SYNC();
    }
    SFB_TILE_SYNC();
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
#include "sfb_step_kernel_r.cuh"
#else

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    const size_t fixed = (size_t)2 * kNF * kTN * 16 + (size_t)kNSc * kTN * 8 + 16 + kRingBytes;
    const size_t per_buf = (size_t)kNRow * kTN * 16;
    const size_t lim = 227 * 1024;
    const int nbuf_rk = kNBufRK;
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
#ifdef SFB_LOOP
    { void* tp = nullptr; e = cudaGetSymbolAddress(&tp, sfb_ltab); if (e != cudaSuccess) return e; P.ktab = reinterpret_cast<const double2*>(tp); }
#endif
#ifdef SFB_GTAB
    { void* tp = nullptr; e = cudaGetSymbolAddress(&tp, sfb_gtab); if (e != cudaSuccess) return e; P.ktab = reinterpret_cast<const double2*>(tp); }
#endif
    P.n0_global = 1;
    const int nbuf = P.nstage == 1 ? 1 : nbuf_rk;
    const size_t smem = nbuf * per_buf + fixed;
    if (smem > lim) return cudaErrorInvalidConfiguration;
    {   // regularisation constants live in this unit's __constant__ bank; refresh when they change
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
