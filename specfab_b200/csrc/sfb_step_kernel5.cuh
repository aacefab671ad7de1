// PERSISTENT fused fabric-evolution step kernel (sm_100a): four lanes per node, one lock-stepped
// instruction stream per SM, streaming TMA refill.  One translation unit per (L, term set).
//
//   d nlm/dt = (M_LROT + Gamma0*M_DDRX + Lambda*M_CDRX + M_REG) nlm      Euler / RK4
//   (reference: src/dynamics.f90:52-97, 251-298, 402-422, 474-518, 108; RK4 per BASELINE config 2)
//
// Why this shape (profiles/r01_notes.md): the generated straight-line operator code is 30 KB (L=8) to
// 0.5 MB (L=20) per stage; with several independent CTAs per SM the warps sit at unrelated positions
// of it and stall on instruction fetch (`no_instruction` 3-8 warps per issue at L>=12).  Here ONE CTA
// per SM runs all its warps through the same code in lock step (a barrier per canonical-m block), so
// the stream is fetched once per tile.  The per-tile load bubble that one-CTA-per-SM would expose is
// removed without a second state buffer: in the LAST stage of a tile the rows of the input buffer die
// in canonical-m order, and right after the barrier that retires a row the next tile's row is bulk-
// copied (TMA, mbarrier complete_tx) into the same place.  The next tile's velocity gradients /
// stresses are staged the same way into a small double-buffered area.
//
// Included by generated .cu files that define:
//   SFB_L, SFB_DDRX (0/1), SFB_TN, SFB_MINB, SFB_WINDOW (0/1), SFB_NAME (launcher symbol), SFB_APPLY_INC
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
constexpr int kNSc = SFB_DDRX ? 17 : 4;                   // per-node scalars
constexpr int kNRaw = SFB_DDRX ? 18 : 9;                  // staged forcing planes: ugrad (9) [+ tau (9)]
constexpr int kDmax = SFB_DDRX ? 4 : 2;                   // |m_i - m_j| range of the operators
static_assert(kTN % 8 == 0, "tile must be a multiple of 8 nodes (one warp)");

__constant__ SfbRegConst c_reg;

__host__ __device__ constexpr int pslot(int l, int a) { return (l / 2) * (l / 2) + a; }
__host__ __device__ constexpr int hrow(int l) { return l * (l + 1) / 2; }

enum { SC_C0 = 0, SC_LAM = 1, SC_RM = 2, SC_G0 = 3, SC_TAUV = 4, SC_TSQV = 10, SC_NORM = 16 };

#define SFB_HORNER (!SFB_DDRX)     // RK4 formulation, see sfb_step_kernel4.cuh

struct Refill {                    // state of the streaming refill (meaningful in thread 0 only)
    const double2* src;            // next tile: P.nlm_in + next_node0
    long long ld;
    uint32_t dst;                  // shared address of the buffer being retired
    uint32_t bytes;                // bytes per row = nvalid_next * 16
    uint32_t mbar;
};

struct Ctx {
    const double *yz, *yp, *yn;
    const double2* fz;
    double *oz, *op;
    double *az, *ap;
    double* gout;
    const double* gin;
    long long ld_out, sld, ld_in, sld_in;
    double c0, lam, rm;
    double as, bs;
    double sigma;
    bool first, last, isA, valid, ld_n0, ld_acc, refill;
    Refill rf;
};

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
    if (c.ld_acc) v = (mu == 0 ? c.az : c.ap)[4 * pslot(l, mu) * kTN];
    return v;
#endif
}
template <int l, int mu>
__device__ __forceinline__ void row_out(const Ctx& c, double mine, double theirs, double z, double n0, double acc) {
    const double recv = __shfl_xor_sync(0xffffffffu, theirs, 1);
    double k = fma(c.sigma, recv, mine);
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    k = fma(d, z, k);
    constexpr int off = 4 * pslot(l, mu) * kTN;
    const bool own = (mu != 0) || c.isA;
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

// bulk-copy the next tile's rows (l, +-a), l = a.., into the retired slots a of the buffer
template <int a>
__device__ __forceinline__ void refill_slot(const Refill& r) {
    constexpr int l0 = a + (a & 1);
#pragma unroll
    for (int l = l0; l <= kL; l += 2) {
#pragma unroll
        for (int sgn = 0; sgn < (a == 0 ? 1 : 2); ++sgn) {
            const int j = hrow(l) + (sgn ? -a : a);
            const uint32_t dst = r.dst + (uint32_t)((2 * pslot(l, a) + sgn) * kTN * 16);
            const double2* src = r.src + (long long)j * r.ld;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(r.bytes), "r"(r.mbar) : "memory");
        }
    }
}
template <int a, int a1>
__device__ __forceinline__ void refill_slots(const Refill& r) {
    if constexpr (a <= a1 && a <= kL) {
        if constexpr (a >= 0) refill_slot<a>(r);
        refill_slots<a + 1, a1>(r);
    }
}
// after the barrier that follows canonical blocks lo..hi of the last stage
//   register window : every value of column nu is read at its first use, block max(0, nu - Dmax)
//                     -> after blocks lo..hi the slots  (lo == 0 ? 0 : lo + Dmax) .. hi + Dmax  are dead
//   no window       : column nu is read up to block nu + Dmax -> slots lo - Dmax .. hi - Dmax are dead,
//                     and everything left once the last block (hi == L) is done
template <int lo, int hi>
__device__ __forceinline__ void refill_after(const Refill& r) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#if SFB_WINDOW
    refill_slots<(lo == 0 ? 0 : lo + kDmax), hi + kDmax>(r);
#else
    refill_slots<lo - kDmax, (hi == kL ? kL : hi - kDmax)>(r);
#endif
}

#define SFB_ROW_OUT4(l, mu, m, t, z, q, r) row_out<l, mu>(c, m, t, z, q, r)
#define SFB_N0_LOAD(l, mu) n0_load<l, mu>(c)
#define SFB_ACC_LOAD(l, mu) acc_load<l, mu>(c)
#define SFB_LOCKSTEP(lo, hi)                          \
    do {                                              \
        __syncthreads();                              \
        if (c.refill) refill_after<lo, hi>(c.rf);     \
    } while (0)

__device__ __forceinline__ void apply_all(const Ctx& c) {
    const double* __restrict__ yz = c.yz;
    const double* __restrict__ yp = c.yp;
    const double* __restrict__ yn = c.yn;
    const double2* __restrict__ fz = c.fz;
#include SFB_APPLY_INC
}

#include "sfb_step_common.cuh"

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(kThreads, 1) step_kernel(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nbuf = P.nstage == 1 ? 1 : (SFB_HORNER ? 2 : 3);
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRow * kTN;
    double* scal = reinterpret_cast<double*>(forc + 2 * kNF * kTN);
    double* raw = scal + kNSc * kTN;                                  // [2][kNRaw][kTN]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(raw + 2 * kNRaw * kTN);   // [0] state, [1..2] raw

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sb = lane >> 4, comp = lane & 1;
    const int nl = warp * 8 + ((lane & 15) >> 1);
    const long long ntiles = (P.N + kTN - 1) / kTN;
    long long tile = blockIdx.x;
    if (tile >= ntiles) return;

    const uint32_t mb_state = smem_u32(mbar), mb_raw0 = smem_u32(mbar + 1), mb_raw1 = smem_u32(mbar + 2);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_state));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_raw0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_raw1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_raw = [&](long long t_node0, int nv, int rb) {      // thread 0: stage ugrad (+tau) of a tile
        const uint32_t mb = rb ? mb_raw1 : mb_raw0;
        const uint32_t bytes = (uint32_t)nv * 8u;
        const int nplanes = (SFB_DDRX && P.tau) ? 18 : 9;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes * (uint32_t)nplanes) : "memory");
        for (int q = 0; q < nplanes; ++q) {
            const double* src = (q < 9 ? P.ugrad + (long long)q * P.ld_u : P.tau + (long long)(q - 9) * P.ld_t) + t_node0;
            const uint32_t dst = smem_u32(raw + ((size_t)rb * kNRaw + q) * kTN);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    };

    // ---- first tile: everything at once
    {
        const long long node0 = tile * kTN;
        const int nv = (int)min((long long)kTN, P.N - node0);
        if (warp == 0) {
            const uint32_t bytes = (uint32_t)nv * 16u;
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_state), "r"(bytes * (uint32_t)kNCoef) : "memory");
            for (int j = lane; j < kNCoef; j += 32) {
                int l = 0;
                while ((l + 2) * (l + 3) / 2 - (l + 2) <= j) l += 2;
                const int m = j - hrow(l);
                const int prow = 2 * pslot(l, m < 0 ? -m : m) + (m < 0 ? 1 : 0);
                const uint32_t dst = smem_u32(bufs + (size_t)prow * kTN);
                const double2* src = P.nlm_in + (long long)j * P.ld_in + node0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src), "r"(bytes), "r"(mb_state) : "memory");
            }
        }
        if (tid == 0 && P.raw_ok && (nv % 2 == 0)) issue_raw(node0, nv, 0);
    }

    Ctx c;
    c.isA = (sb == 0);
    c.sigma = comp ? 1.0 : -1.0;
    c.fz = forc + (size_t)sb * kNF * kTN + nl;
    c.ld_out = P.ld_out; c.sld = sb ? -P.ld_out : P.ld_out;
    c.ld_in = P.ld_in;   c.sld_in = sb ? -P.ld_in : P.ld_in;
    c.rf.ld = P.ld_in;
    c.rf.mbar = mb_state;
    double* dbuf = reinterpret_cast<double*>(bufs) + 2 * nl + comp;
    constexpr size_t kBufD = (size_t)2 * kNRow * kTN;
    constexpr int kPlane = 2 * kTN;
    double* acc = dbuf + (size_t)(nbuf - 1) * kBufD;
    c.az = acc; c.ap = acc + sb * kPlane;

    int ib0 = 0;
    uint32_t rawphase[2] = {0u, 0u};
    for (unsigned it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const long long node0 = tile * kTN;
        const int nvalid = (int)min((long long)kTN, P.N - node0);
        const long long next = tile + gridDim.x;
        const bool has_next = next < ntiles;
        const int rb = it & 1;
        c.valid = nl < nvalid;
        c.gout = reinterpret_cast<double*>(P.nlm_out + node0 + nl) + comp;
        c.gin = reinterpret_cast<const double*>(P.nlm_in + node0 + nl) + comp;

        // ---- forcing of this tile (task-parallel over the CTA's threads) from the TMA-staged planes; tiles whose
        //      planes cannot be bulk-copied (16-byte alignment / even node count) read global memory directly
        const bool use_raw = P.raw_ok && (nvalid % 2 == 0);
        if (use_raw) {
            mbar_wait(rb ? mb_raw1 : mb_raw0, rawphase[rb]);
            rawphase[rb] ^= 1;
        }
        {
            const int task = tid / kTN, t = tid - task * kTN;
            if (t < nvalid) {
                const double* rp = raw + (size_t)rb * kNRaw * kTN + t;
                ForcSrc S;
                if (use_raw) {
                    S.ug = rp; S.su = kTN;
                    S.tau = (SFB_DDRX && P.tau) ? rp + 9 * kTN : nullptr; S.st = kTN;
                } else {
                    S = global_src(P, node0 + t);
                }
                if (task == 0) prep_lrot(P, S, node0 + t, t, forc, scal);
#if SFB_DDRX
                if (task == 1) prep_ddrx_g(P, S, node0 + t, t, forc, scal);
                if (task == 2) prep_ddrx_d(S, t, scal);
#endif
            }
        }
        mbar_wait(mb_state, it & 1);
        __syncthreads();
        if (has_next && tid == 0) {
            const long long nn0 = next * kTN;
            const int nvn = (int)min((long long)kTN, P.N - nn0);
            if (P.raw_ok && (nvn % 2 == 0)) issue_raw(nn0, nvn, rb ^ 1);
            // arm the state barrier for the next tile; the copies are issued while the last stage retires rows
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_state), "r"((uint32_t)nvn * 16u * (uint32_t)kNCoef) : "memory");
            c.rf.src = P.nlm_in + nn0;
            c.rf.bytes = (uint32_t)nvn * 16u;
        }
        c.lam = scal[SC_LAM * kTN + nl];
        c.rm = scal[SC_RM * kTN + nl];
        c.c0 = 0.0;

        for (int s = 0; s < P.nstage; ++s) {
            const int ib = (nbuf == 1) ? 0 : ((ib0 + s) & 1), ob = ib ^ 1;
            const double* yin = dbuf + (size_t)ib * kBufD;
            double* yout = dbuf + (size_t)ob * kBufD;
            c.yz = yin; c.yp = yin + sb * kPlane; c.yn = yin + (1 - sb) * kPlane;
            c.oz = yout; c.op = yout + sb * kPlane;
            c.first = (s == 0);
            c.last = (s == P.nstage - 1);
            c.ld_n0 = !c.first && c.valid && (SFB_HORNER || !c.last);
            c.ld_acc = !c.first;
            c.refill = c.last && has_next && tid == 0;
            c.rf.dst = smem_u32(bufs + (size_t)ib * kNRow * kTN);
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
                double tv[6], sv[6];
#pragma unroll
                for (int p = 0; p < 6; ++p) { tv[p] = scal[(SC_TAUV + p) * kTN + tid]; sv[p] = scal[(SC_TSQV + p) * kTN + tid]; }
                const double davg = sfb::ev_D2(y[0], n2, n4, tv, sv, scal[SC_NORM * kTN + tid]);
                scal[SC_C0 * kTN + tid] = -(scal[SC_G0 * kTN + tid] * davg);
            }
            __syncthreads();
            c.c0 = scal[SC_C0 * kTN + nl];
#endif
            apply_all(c);
            __syncthreads();
        }
        if (nbuf > 1) ib0 = (ib0 + P.nstage - 1) & 1;
    }
}

}  // namespace

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    static int grid_max[64] = {0};
    const size_t fixed = (size_t)2 * kNF * kTN * 16 + (size_t)kNSc * kTN * 8 + (size_t)2 * kNRaw * kTN * 8 + 32;
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
        int nsm = 0;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        grid_max[dev] = nsm > 0 ? nsm : 148;
        attr_done[dev] = true;
    }
    attr_lk.unlock();
    if (Pin.rio) return cudaErrorNotSupported;      // reduced-form arrays: reduced kernels only
    SfbStepParams P = Pin;
    P.n0_global = 1;
    P.raw_ok = (((uintptr_t)P.ugrad & 15) == 0 && (P.ld_u % 2) == 0 &&
                (!P.tau || (((uintptr_t)P.tau & 15) == 0 && (P.ld_t % 2) == 0))) ? 1 : 0;
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
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel, kThreads, smem);
    if (occ < 1) occ = 1;
    const long long grid = ntile < (long long)grid_max[dev] * occ ? ntile : (long long)grid_max[dev] * occ;
    step_kernel<<<(unsigned)grid, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}
