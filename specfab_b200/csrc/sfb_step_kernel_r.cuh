// Reduced-form fused step kernel: ONE lane per node for states with the real-ODF symmetry
//     n_l^{-m} = (-1)^m conj(n_l^m),  Im n_l^0 = 0            (every physical ODF; src/reducedform.f90:160-170).
// M_LROT, M_DDRX_src and the diagonal terms commute with this mirror (exact mirror symmetry of the Gaunt tables,
// codegen/operators.py), so the rows m < 0 of the new state are the mirror of the rows m >= 0 and only the
// latter are computed: half the DFMAs, half the shared memory and half of the state read of the full kernel.
// The mirror rows are still WRITTEN (the output is a full nlm array, like the reference's).
//
// The symmetry is verified per tile against the rows m < 0 of the input (read once from global, never staged),
// to round-off: every component of n_l^{-m} - (-1)^m conj(n_l^m) and of Im n_l^0 must be within kSymTol * |n_0^0|
// (2^-46 ~ 1.4e-14: what a chain of FP64 steps of the full algorithm -- or of the reference -- leaves behind;
// two orders below the 1e-12 per-step parity bar).  The reduced path then works from the rows m >= 0 and writes an
// exactly symmetric state.  A tile that fails the test -- general complex vectors are legal input for the
// reference operators -- is handed to full_tile() (sfb_step_kernel.cuh), the unreduced two-lane algorithm, 16
// nodes at a time inside the same CTA and the same shared-memory footprint.  Included at the end of sfb_step_kernel.cuh
// when SFB_REDUCED is defined; SFB_TNR = nodes per CTA here (= 2 * SFB_TN), SFB_APPLY_INC_R = generated body.
#pragma once

namespace {

constexpr int kTNR = SFB_TNR;
// SFB_N0ALL (Horner kernels): n0 is loaded from global in EVERY stage -- no select between the stage input and the loaded
// value, no zero initialisation, no branch around the loads; stage 0 pays 16 B per row of L2 (not DRAM) traffic for it
#if defined(SFB_N0ALL_REQ) && SFB_HORNER
#define SFB_N0ALL 1
#else
#define SFB_N0ALL 0
#endif
constexpr double kSymTol = 0x1p-46;
constexpr int kNRowR = (kL / 2 + 1) * (kL / 2 + 1);       // rows (l, m >= 0)
// one lane per node; kR warp roles share the 32 nodes of the tile (straight-line form: kR = 1, no barriers at all;
// table-driven loop form: the (mu, chunk) items are dealt to the roles).  The fallback runs full_tile() with the same
// kThreads = 32 * kR threads on tiles of 16 nodes.
static_assert(kTNR == 32 && kTN == 16 && kThreads == 32 * kR, "reduced kernel: 32 nodes per CTA, fallback tiles of 16");

struct CtxR {
    const double2* ktab;              // loop mode: global table of operator entries
    double2* ring;                    // loop mode: this warp's private table ring in shared memory
    const double2* yp;                // stage input, rows pslot(l, m) * kTNR
    const double2* fz;                // forcing block (lane set A's)
    double2* op;                      // next-stage buffer
    double2* ap;                      // RK accumulator buffer (classical RK4)
    double2* gout;                    // global output, offset by node
    const double2* gin;
    long long ld_out, ld_in;
    unsigned ldo32, ldi32;            // SFB_A32: the same strides as 32-bit values
    double c0, lam, rm;
    double as, bs;
    bool first, last, valid, ld_n0, ld_acc;
};

// Global row addresses of the row finalisation.  SFB_A32: base + (16 * row) * ld as ONE 32x32->64 multiply-add
// (IMAD.WIDE.U32; the launcher rejects ld >= 2^32) and explicit global-space accesses (LDG/STG, L2-only loads, streaming
// stores) instead of a 64-bit multiply (4-5 integer instructions per address) and generic LD/ST.
#ifdef SFB_A32
__device__ __forceinline__ unsigned long long mad_wide(unsigned a, unsigned b, unsigned long long c) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));     // opaque to the optimiser: stays one instruction
    return r;
}
template <int row>
__device__ __forceinline__ const double2* grow_in(const CtxR& c) {
    return reinterpret_cast<const double2*>(mad_wide(c.ldi32, 16u * row, reinterpret_cast<unsigned long long>(c.gin)));
}
template <int row>
__device__ __forceinline__ double2* grow_out(const CtxR& c) {
    return reinterpret_cast<double2*>(mad_wide(c.ldo32, 16u * row, reinterpret_cast<unsigned long long>(c.gout)));
}
__device__ __forceinline__ double2 gload(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ void gstore(double2* p, double2 v) { __stcs(p, v); }
#else
template <int row>
__device__ __forceinline__ const double2* grow_in(const CtxR& c) { return c.gin + (long long)row * c.ld_in; }
template <int row>
__device__ __forceinline__ double2* grow_out(const CtxR& c) { return c.gout + (long long)row * c.ld_out; }
__device__ __forceinline__ double2 gload(const double2* p) { return *p; }
__device__ __forceinline__ void gstore(double2* p, double2 v) { *p = v; }
#endif

// RIO ("reduced I/O", sfb_step_rnlm_arr): the global arrays hold the rows m >= 0 only, row (l, m) at (l/2)^2 + m -- the
// rnlm layout of src/reducedform.f90:160-187 -- so nothing is tested, no mirror row is read or written
template <int l, int mu, bool RIO>
__host__ __device__ constexpr int grow_idx() { return RIO ? pslot(l, mu) : hrow(l) + mu; }

template <int l, int mu, bool RIO>
__device__ __forceinline__ double2 n0_load_r(const CtxR& c) {
#if SFB_N0ALL
    return gload(grow_in<grow_idx<l, mu, RIO>()>(c));      // every stage, every lane (c.gin of a lane beyond N points at the tile's first node)
#else
    double2 v = make_double2(0.0, 0.0);
    if (c.ld_n0) v = gload(grow_in<grow_idx<l, mu, RIO>()>(c));
    return v;
#endif
}
template <int l, int mu>
__device__ __forceinline__ double2 acc_load_r(const CtxR& c) {
    double2 v = make_double2(0.0, 0.0);
#if !SFB_HORNER
    if (c.ld_acc) v = c.ap[pslot(l, mu) * kTNR];
#endif
    return v;
}
// finalize one row; returns the stage output y.  STORE: write y to the next-stage buffer now (two-buffer scheme);
// otherwise the caller parks it in a register and commits it later (in-place scheme, SFB_INPLACE)
template <int l, int mu, bool STORE, bool RIO>
__device__ __forceinline__ double2 row_out_r(const CtxR& c, double kr, double ki, double zr, double zi, double2 n0, double2 acc) {
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    kr = fma(d, zr, kr);
    ki = fma(d, zi, ki);
    if (mu == 0) { ki = 0.0; zi = 0.0; n0.y = 0.0; acc.y = 0.0; }     // n_l^0 of a real ODF is real: drop the round-off
    constexpr int off = pslot(l, mu) * kTNR;
#if SFB_N0ALL
    const double n0r = n0.x, n0i = n0.y;         // stage 0 re-reads the rows the bulk copies have just pulled through L2
#else
    const double n0r = c.first ? zr : n0.x, n0i = c.first ? zi : n0.y;
#endif
#if SFB_HORNER
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (STORE && !c.last) c.op[off] = y;
    const double2 res = y;
#else
    const double2 A = make_double2(fma(c.bs, kr, c.first ? zr : acc.x), fma(c.bs, ki, c.first ? zi : acc.y));
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (!c.last) { if (STORE) c.op[off] = y; c.ap[off] = A; }
    const double2 res = A;
#endif
    if (c.last && c.valid) {
        gstore(grow_out<grow_idx<l, mu, RIO>()>(c), res);
        if (mu != 0 && !RIO)      // mirror row: (-1)^mu conj
            gstore(grow_out<hrow(l) - mu>(c), (mu & 1) ? make_double2(-res.x, res.y) : make_double2(res.x, -res.y));
    }
    return y;
}
// (RIO: template parameter of the enclosing apply_reduced)
#define SFB_RROW_PRE(l, mu, q, r) const double2 q = n0_load_r<l, mu, RIO>(c), r = acc_load_r<l, mu>(c)
#define SFB_RROW_OUT(l, mu, ar, ai, zr, zi, q, r) row_out_r<l, mu, true, RIO>(c, ar, ai, zr, zi, q, r)
#define SFB_RROW_OUTQ(l, mu, ar, ai, zr, zi, q, r, o) o = row_out_r<l, mu, false, RIO>(c, ar, ai, zr, zi, q, r)
#define SFB_RROW_COMMIT(l, mu, o) do { if (!c.last) c.op[pslot(l, mu) * kTNR] = o; } while (0)

#ifdef SFB_LOOP
#include "sfb_step_loop_r.cuh"
template <bool RIO>
__device__ __forceinline__ void apply_reduced(const CtxR& c, int role) { loopk::apply_loop_r<RIO>(c, role, kR, c.ring, (int)(threadIdx.x & 31)); }
#else
template <bool RIO>
__device__ __forceinline__ void apply_reduced(const CtxR& c, int role) {
    const double2* __restrict__ yp = c.yp;
    const double2* __restrict__ fz = c.fz;
#include SFB_APPLY_INC_R
}
#endif

// bytes of one tile's shared-memory slice (stage buffers | forcing | scalars | mbarrier | table rings); 128-byte granular for kCW > 1
__host__ __device__ constexpr size_t slice_bytes(int nbuf) {
    const size_t b = (size_t)nbuf * kNRowR * kTNR * 16 + (size_t)kNF * kTNR * 16 + (size_t)kNSc * kTNR * 8 + 16 + kRingBytes;
    return kCW > 1 ? (b + 127) / 128 * 128 : b;
}

template <bool RIO>
__device__ __forceinline__ void step_tile_r(const SfbStepParams& P, unsigned char* smem_cta) {
    const int nbuf = P.nstage == 1 ? 1 : kNBufRK;
    const int sub = threadIdx.x / kThreads;                 // tile of this CTA (kCW > 1: one warp each)
    unsigned char* smem_raw = smem_cta + (kCW > 1 ? (size_t)sub * slice_bytes(nbuf) : 0);
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRowR * kTNR;
    double* scal = reinterpret_cast<double*>(forc + kNF * kTNR);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(scal + kNSc * kTNR);
    double2* rings = reinterpret_cast<double2*>(mbar + 2);          // loop mode: [warps][2][pairs per item]
    (void)rings;

    const int tid = threadIdx.x % kThreads, warp = tid >> 5;
    const int t = tid & 31;                           // node within tile
    const long long node0 = ((long long)blockIdx.x * kCW + sub) * kTNR;
#ifdef SFB_LS
    constexpr bool kLS = kCW > 1;                     // lock step: the general-state decision is CTA wide, the reduced body has CTA barriers
#else
    constexpr bool kLS = false;
#endif
    if (kCW > 1 && node0 >= P.N) return;              // exited warps do not count at later CTA barriers
    const int nvalid = (int)min((long long)kTNR, P.N - node0);
    const bool valid = t < nvalid;
    // the node's velocity gradient, needed first thing by the forcing preparation: requested before the tile is staged so that
    // its DRAM latency runs under the mbarrier set-up and the bulk-copy issue loop
    double upre[9];
    if (valid && warp == 0) {
#pragma unroll
        for (int p = 0; p < 9; ++p) upre[p] = P.ugrad[(long long)p * P.ld_u + node0 + t];
    }

    // ---- stage the rows m >= 0 of the tile (one bulk copy per row) into buffer 0
    const uint32_t mb = smem_u32(mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    SFB_TILE_SYNC();
    if (warp == 0) {
        const uint32_t bytes = (uint32_t)nvalid * 16u;
        if (t == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes * (uint32_t)kNRowR) : "memory");
        __syncwarp();
        for (int r = t; r < kNRowR; r += 32) {
            // reduced row r = (l/2)^2 + m
            int h = 0;
            while ((h + 1) * (h + 1) <= r) ++h;
            const int l = 2 * h, m = r - h * h;
            const uint32_t dst = smem_u32(bufs + (size_t)r * kTNR);
            const double2* src = P.nlm_in + (long long)(RIO ? r : hrow(l) + m) * P.ld_in + node0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    }
    // ---- the mirror rows (m < 0) of this lane's node, for the symmetry test (warp 0): issued now, all loads in flight
    // together, consumed after the tile has landed (their latency hides behind the bulk copies and the forcing preparation)
    constexpr int kNNeg = kNCoef - kNRowR;
    constexpr bool kBatchAll = kNNeg <= 42 && !SFB_DDRX && !RIO;   // LROT kernels up to L = 12: every mirror row in registers
                                                         // (the DDRX preparation needs the registers: per-degree batches there)
    if (!RIO && !kBatchAll && warp == (kR > 1 ? kR - 1 : 0))
        prefetch_mirror_rows(P.nlm_in, P.ld_in, node0, nvalid, kL, t, 32);        // tested after the wait: have them in L2 by then
    double2 vneg[kBatchAll ? (kNNeg > 0 ? kNNeg : 1) : 1];
    const double2* gneg = P.nlm_in + node0 + (valid ? t : 0);
    if (!RIO && kBatchAll && warp == 0) {
#pragma unroll
        for (int l = 2; l <= kL; l += 2)
#pragma unroll
            for (int m = 1; m <= l; ++m) vneg[(l / 2) * (l / 2 - 1) + m - 1] = gneg[(long long)(hrow(l) - m) * P.ld_in];
    }
    // ---- meanwhile: per-node forcing (lane set A only); with several roles the tasks go to different warps
    if (valid) {
        const ForcSrc S = global_src(P, node0 + t);
        if (warp == 0) prep_lrot<kTNR, false>(P, S, node0 + t, t, forc, scal, upre);
#if SFB_DDRX
        if (warp == (kR > 1 ? 1 : 0)) prep_ddrx_g<kTNR, false>(P, S, node0 + t, t, forc, scal);
#endif
    }
    {   // wait for the tile
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }

    // ---- real-ODF symmetry of the input to round-off (NaNs fail the test and take the general path)
    // (with several roles the degrees l are dealt to the warps; a single warp owns all of them and may have batched the loads)
    bool bad = false;
    if (!RIO) {
        const double tol = kSymTol * fabs(bufs[t].x);
#pragma unroll
        for (int l = 0; l <= kL; l += 2) {
            if (kBatchAll ? (warp != 0) : (((l / 2) % kR) != warp)) continue;
            bad |= !(fabs(bufs[(size_t)pslot(l, 0) * kTNR + t].y) <= tol);
            double2 vl[kL > 0 ? kL : 1];
            if (!kBatchAll) {
#pragma unroll
                for (int m = 1; m <= kL; ++m)
                    if (m <= l) vl[m - 1] = gneg[(long long)(hrow(l) - m) * P.ld_in];
            }
#pragma unroll
            for (int m = 1; m <= kL; ++m)
                if (m <= l) {
                    const double2 vp = bufs[(size_t)pslot(l, m) * kTNR + t];
                    const double2 vn = kBatchAll ? vneg[(l / 2) * (l / 2 - 1) + m - 1] : vl[m - 1];
                    const double er = (m & 1) ? -vp.x : vp.x, ei = (m & 1) ? vp.y : -vp.y;
                    bad |= !(fabs(vn.x - er) <= tol && fabs(vn.y - ei) <= tol);
                }
        }
        bad = bad && valid;
    }
    if (RIO) {                          // a reduced-form state is symmetric by construction: only order the forcing preparation
        if (kR > 1) SFB_TILE_SYNC(); else __syncwarp();
    } else if ((kCW > 1 && !kLS) ? (__any_sync(0xffffffffu, bad) != 0) : (__syncthreads_or(bad) != 0)) {     // also orders the forcing preparation before the stages
        if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");
        SFB_TILE_SYNC();
#ifdef SFB_NOFB
        __trap();                       // experiment only: how much does the presence of the fallback call cost the reduced path?
#else
        full_tile(P, node0, smem_raw);
        if (node0 + kTN < P.N) full_tile(P, node0 + kTN, smem_raw);
#endif
        return;
    }

    CtxR c;
#ifdef SFB_LOOP
    c.ring = rings + (size_t)warp * 2 * (SFB_LT_PER_CHUNK / 2);
    c.ktab = P.ktab;
#else
    c.ring = nullptr;
    c.ktab = nullptr;
#endif
    c.valid = valid;
    c.fz = forc + t;
    c.lam = scal[SC_LAM * kTNR + t];
    c.rm = scal[SC_RM * kTNR + t];
    c.c0 = 0.0;
    c.ld_out = P.ld_out;
    c.ld_in = P.ld_in;
    c.ldo32 = (unsigned)P.ld_out;
    c.ldi32 = (unsigned)P.ld_in;
    c.gout = P.nlm_out + node0 + t;
    c.gin = P.nlm_in + node0 + (SFB_N0ALL && !valid ? 0 : t);
    c.ap = bufs + (size_t)(nbuf - 1) * kNRowR * kTNR + t;

    for (int s = 0; s < P.nstage; ++s) {
        const int ib = kInPlace ? 0 : (s & 1), ob = kInPlace ? 0 : ((s + 1) & 1);
        c.yp = bufs + (size_t)ib * kNRowR * kTNR + t;
        c.op = bufs + (size_t)ob * kNRowR * kTNR + t;
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
        if (kR == 1 || warp == 0) {   // <D>(current stage state), one lane per node
            if (valid) {
                const double2* y = c.yp;
                double2 n2[3], n4[5];
#pragma unroll
                for (int m = 0; m < 3; ++m) n2[m] = y[pslot(2, m) * kTNR];
#pragma unroll
                for (int m = 0; m < 5; ++m) n4[m] = (kL >= 4) ? y[pslot(4, m) * kTNR] : make_double2(0.0, 0.0);
                const double davg = ddrx_davg(global_src(P, node0 + t), y[0], n2, n4);
                c.c0 = -(scal[SC_G0 * kTNR + t] * davg);
                if (kR > 1) scal[SC_C0 * kTNR + t] = c.c0;
            }
        }
        if (kR > 1) {
            SFB_TILE_SYNC();
            c.c0 = scal[SC_C0 * kTNR + t];
        }
#endif
        apply_reduced<RIO>(c, warp);
        // one role: every lane reads and writes only its own node's column -- no barrier between stages
        if (kR > 1 && !c.last) SFB_TILE_SYNC();
    }
}

__global__ void __launch_bounds__(kThreads * kCW, SFB_MINB) step_kernel_r(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_cta[];
    step_tile_r<false>(P, smem_cta);
}
// the same step on reduced-form arrays (rows m >= 0 in, rows m >= 0 out)
__global__ void __launch_bounds__(kThreads * kCW, SFB_MINB) step_kernel_rr(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_cta[];
    step_tile_r<true>(P, smem_cta);
}

}  // namespace

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    const int nbuf_rk = kNBufRK;
    const size_t smem_max = slice_bytes(nbuf_rk) * kCW;
    // the general-path fallback (full_tile, 16 nodes, both row planes, both forcing blocks) uses the same bytes
    static_assert((size_t)kNRow * kTN == (size_t)kNRowR * kTNR && 2 * kNF * kTN == kNF * kTNR, "fallback must fit the reduced layout");
    if (smem_max > 227 * 1024) return cudaErrorInvalidConfiguration;
#ifdef SFB_A32
    if (Pin.ld_in >= (1LL << 32) || Pin.ld_out >= (1LL << 32)) return cudaErrorInvalidValue;   // 32-bit row strides
#endif
    cudaError_t e;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::unique_lock<std::mutex> attr_lk(attr_mu);
    if (!attr_done[dev]) {
        e = cudaFuncSetAttribute(step_kernel_r, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(step_kernel_rr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    attr_lk.unlock();
    SfbStepParams P = Pin;
#ifdef SFB_LOOP
    { void* tp = nullptr; e = cudaGetSymbolAddress(&tp, sfb_ltab); if (e != cudaSuccess) return e; P.ktab = reinterpret_cast<const double2*>(tp); }
#endif
    P.n0_global = 1;
    const int nbuf = P.nstage == 1 ? 1 : nbuf_rk;
    const size_t smem = slice_bytes(nbuf) * kCW;
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
    const long long ntile = (P.N + (long long)kTNR * kCW - 1) / ((long long)kTNR * kCW);
    if (P.rio) step_kernel_rr<<<(unsigned)ntile, kThreads * kCW, smem, st>>>(P);
    else step_kernel_r<<<(unsigned)ntile, kThreads * kCW, smem, st>>>(P);
    return cudaGetLastError();
}
