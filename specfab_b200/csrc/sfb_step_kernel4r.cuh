// Reduced form of the four-lane step kernel: TWO lanes per node (re | im) for states with the real-ODF symmetry
//     n_l^{-m} = (-1)^m conj(n_l^m),  Im n_l^0 = 0.
// Same idea as sfb_step_kernel_r.cuh (read that header first): only the rows m >= 0 are staged and computed, a
// column block with nu < 0 is read from the rows (l_j, |nu|); in the re|im lane split the conj-mirror of such a block,
// S = (-1)^nu conj(S'), amounts to scaling its forcing by (-1)^nu on the re lane and -(-1)^nu on the im lane
// (k.x = f.x sr - f.y si, k.y = f.x si + f.y sr with sr = s sr', si = -s si'; the partner exchange keeps its signs).
// A warp holds 16 nodes; a CTA SFB_TNR nodes = 2*SFB_TNR threads in lock step.  Tiles that fail the symmetry test are
// handed to full_tile() of sfb_step_kernel4.cuh (four lanes per node, SFB_TN = SFB_TNR/2 nodes at a time).
// Included at the end of sfb_step_kernel4.cuh when SFB_REDUCED is defined.
#pragma once

namespace {

constexpr int kTNR = SFB_TNR;
constexpr double kSymTol = 0x1p-46;
constexpr int kNRowR = (kL / 2 + 1) * (kL / 2 + 1);       // rows (l, m >= 0)
static_assert(kTNR == 2 * kTN && kThreads == 2 * kTNR, "reduced four-lane kernel: fallback tiles of SFB_TNR/2 nodes");

struct CtxR {
    const double* yp;                 // stage input, this lane's component: row (l,m) at 2*pslot(l,m)*kTNR doubles
    const double2* fz;                // forcing block (sign set A)
    double* op;                       // next-stage buffer
    double* ap;                       // RK accumulator buffer
    double* gout;                     // global output: component of node
    const double* gin;
    long long ld_out, ld_in;
    unsigned ldo32, ldi32;            // SFB_A32: the same strides as 32-bit values
    double c0, lam, rm;
    double as, bs;
    double sigma;                     // -1 (re lane) / +1 (im lane): sign of the partner's contribution
    double em;                        // +1 (re lane) / -1 (im lane): forcing sign of mirrored column blocks
    bool first, last, valid, ld_n0, ld_acc, isim;
};

// global row addresses of the row finalisation (see sfb_step_kernel_r.cuh): SFB_A32 = one 32x32->64 multiply-add per
// address (the launcher rejects ld >= 2^32) and explicit global-space accesses
#ifdef SFB_A32
__device__ __forceinline__ unsigned long long mad_wide(unsigned a, unsigned b, unsigned long long c) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
template <int row>
__device__ __forceinline__ const double* grow_in(const CtxR& c) {
    return reinterpret_cast<const double*>(mad_wide(c.ldi32, 16u * row, reinterpret_cast<unsigned long long>(c.gin)));
}
template <int row>
__device__ __forceinline__ double* grow_out(const CtxR& c) {
    return reinterpret_cast<double*>(mad_wide(c.ldo32, 16u * row, reinterpret_cast<unsigned long long>(c.gout)));
}
__device__ __forceinline__ double gload(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void gstore(double* p, double v) { __stcs(p, v); }
#else
template <int row>
__device__ __forceinline__ const double* grow_in(const CtxR& c) { return c.gin + 2 * ((long long)row * c.ld_in); }
template <int row>
__device__ __forceinline__ double* grow_out(const CtxR& c) { return c.gout + 2 * ((long long)row * c.ld_out); }
__device__ __forceinline__ double gload(const double* p) { return *p; }
__device__ __forceinline__ void gstore(double* p, double v) { *p = v; }
#endif

// RIO ("reduced I/O", sfb_step_rnlm_arr): the global arrays hold the rows m >= 0 only, row (l, m) at (l/2)^2 + m
template <int l, int mu, bool RIO>
__host__ __device__ constexpr int grow_idx() { return RIO ? pslot(l, mu) : hrow(l) + mu; }

template <int l, int mu, bool RIO>
__device__ __forceinline__ double n0_load_r(const CtxR& c) {
    double v = 0.0;
    if (c.ld_n0) v = gload(grow_in<grow_idx<l, mu, RIO>()>(c));
    return v;
}
template <int l, int mu>
__device__ __forceinline__ double acc_load_r(const CtxR& c) {
#if SFB_HORNER
    return 0.0;
#else
    double v = 0.0;
    if (c.ld_acc) v = c.ap[2 * pslot(l, mu) * kTNR];
    return v;
#endif
}
template <int l, int mu, bool RIO>
__device__ __forceinline__ void row_out_r(const CtxR& c, double mine, double theirs, double z, double n0, double acc) {
    const double recv = __shfl_xor_sync(0xffffffffu, theirs, 1);
    double k = fma(c.sigma, recv, mine);
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    k = fma(d, z, k);
    if (mu == 0 && c.isim) { k = 0.0; z = 0.0; n0 = 0.0; acc = 0.0; }     // n_l^0 of a real ODF is real
    constexpr int off = 2 * pslot(l, mu) * kTNR;      // doubles
    const double n0v = c.first ? z : n0;
#if SFB_HORNER
    const double y = fma(c.as, k, n0v);
    if (!c.last) c.op[off] = y;
    const double res = y;
#else
    const double A = fma(c.bs, k, c.first ? z : acc);
    const double y = fma(c.as, k, n0v);
    if (!c.last) { c.op[off] = y; c.ap[off] = A; }
    const double res = A;
#endif
    if (c.last && c.valid) {
        gstore(grow_out<grow_idx<l, mu, RIO>()>(c), res);
        // mirror row (-1)^mu conj: the re lane keeps the parity sign, the im lane gets the opposite one
        if (mu != 0 && !RIO) gstore(grow_out<hrow(l) - mu>(c), ((mu & 1) != 0) == c.isim ? res : -res);
    }
}
// (RIO: template parameter of the enclosing apply_reduced)
#define SFB_RROW_OUT4(l, mu, m, t, z, q, r) row_out_r<l, mu, RIO>(c, m, t, z, q, r)
#define SFB_RN0_LOAD(l, mu) n0_load_r<l, mu, RIO>(c)
#define SFB_RACC_LOAD(l, mu) acc_load_r<l, mu>(c)
// forcing of a mirrored column block: sign = (-1)^nu, times +1 / -1 on the re / im lane
#define SFB_FMIR(fidx, sign) make_double2(fz[(fidx) * SFB_TNR].x * ((sign) * c.em), fz[(fidx) * SFB_TNR].y * ((sign) * c.em))

template <bool RIO>
__device__ __forceinline__ void apply_reduced(const CtxR& c) {
    const double* __restrict__ yp = c.yp;
    const double2* __restrict__ fz = c.fz;
#include SFB_APPLY_INC_R
}

template <bool RIO>
__device__ __forceinline__ void step_tile_r(const SfbStepParams& P, unsigned char* smem_raw) {
    const int nbuf = P.nstage == 1 ? 1 : (SFB_HORNER ? 2 : 3);
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRowR * kTNR;
    double* scal = reinterpret_cast<double*>(forc + kNF * kTNR);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(scal + kNSc * kTNR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int comp = lane & 1;                         // 0: real part, 1: imaginary part
    const int nl = warp * 16 + (lane >> 1);            // node within tile
    const long long node0 = (long long)blockIdx.x * kTNR;
    const int nvalid = (int)min((long long)kTNR, P.N - node0);
    const bool valid = nl < nvalid;

    const uint32_t mb = smem_u32(mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!RIO && warp == 1) prefetch_mirror_rows(P.nlm_in, P.ld_in, node0, nvalid, kL, lane, 32);
    if (warp == 0) {
        const uint32_t bytes = (uint32_t)nvalid * 16u;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes * (uint32_t)kNRowR) : "memory");
        __syncwarp();
        for (int r = lane; r < kNRowR; r += 32) {
            int h = 0;
            while ((h + 1) * (h + 1) <= r) ++h;
            const int l = 2 * h, m = r - h * h;
            const uint32_t dst = smem_u32(bufs + (size_t)r * kTNR);
            const double2* src = P.nlm_in + (long long)(RIO ? r : hrow(l) + m) * P.ld_in + node0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    }
    // ---- per-node forcing (sign set A only): two tasks per node, one thread each (kThreads = 2*kTNR)
    for (int w = tid; w < 2 * kTNR; w += kThreads) {
        const int task = w / kTNR, t = w - task * kTNR;
        if (t < nvalid) {
            const ForcSrc S = global_src(P, node0 + t);
            if (task == 0) prep_lrot<kTNR, false>(P, S, node0 + t, t, forc, scal);
#if SFB_DDRX
            if (task == 1) prep_ddrx_g<kTNR, false>(P, S, node0 + t, t, forc, scal);
#endif
        }
    }
    // ---- this lane's component of the mirror rows (m < 0), for the real-ODF test: issued now that the forcing preparation has
    // released its registers, all loads in flight together (L2 hits after the prefetch above), consumed after the wait and <D>
    constexpr int kNNegR = kNCoef - kNRowR;
    double vneg[kNNegR > 0 ? kNNegR : 1];
    if (!RIO && valid) {
        const double* gneg = reinterpret_cast<const double*>(P.nlm_in + node0 + nl) + comp;
#pragma unroll
        for (int l = 2; l <= kL; l += 2)
#pragma unroll
            for (int m = 1; m <= l; ++m) vneg[(l / 2) * (l / 2 - 1) + m - 1] = gneg[2 * ((long long)(hrow(l) - m) * P.ld_in)];
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }
    const double* dbuf = reinterpret_cast<const double*>(bufs) + 2 * nl + comp;      // this lane's component column
#if SFB_DDRX
    // <D> of the first stage needs only the staged state and tau: computed here, in front of the barrier the symmetry verdict
    // needs anyway (the rate factor scal[SC_G0] is another thread's: it is applied after that barrier), so stage 0 has no
    // barrier of its own
    if (tid < nvalid) {
        const double2* y = bufs + tid;
        double2 n2[3], n4[5];
#pragma unroll
        for (int m = 0; m < 3; ++m) n2[m] = y[pslot(2, m) * kTNR];
#pragma unroll
        for (int m = 0; m < 5; ++m) n4[m] = (kL >= 4) ? y[pslot(4, m) * kTNR] : make_double2(0.0, 0.0);
        scal[SC_C0 * kTNR + tid] = ddrx_davg(global_src(P, node0 + tid), y[0], n2, n4);
    }
#endif

    // ---- real-ODF symmetry of the input to round-off, each lane tests its own component of the mirror rows
    bool bad = false;
    if (!RIO && valid) {
        const double tol = kSymTol * fabs(reinterpret_cast<const double*>(bufs)[2 * nl]);
#pragma unroll
        for (int l = 0; l <= kL; l += 2) {
            if (comp) bad |= !(fabs(dbuf[2 * pslot(l, 0) * kTNR]) <= tol);
#pragma unroll
            for (int m = 1; m <= kL; ++m)
                if (m <= l) {
                    const double vp = dbuf[2 * pslot(l, m) * kTNR];
                    const double ex = (((m & 1) != 0) == (comp != 0)) ? vp : -vp;     // re: (-1)^m vp ; im: -(-1)^m vp
                    bad |= !(fabs(vneg[(l / 2) * (l / 2 - 1) + m - 1] - ex) <= tol);
                }
        }
    }
    if (RIO) {                          // a reduced-form state is symmetric by construction: only order the forcing preparation
        __syncthreads();
    } else if (__syncthreads_or(bad)) {
        if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");
        __syncthreads();
        full_tile(P, node0, smem_raw);
        if (node0 + kTN < P.N) full_tile(P, node0 + kTN, smem_raw);
        return;
    }

    CtxR c;
    c.valid = valid;
    c.isim = comp != 0;
    c.sigma = comp ? 1.0 : -1.0;
    c.em = comp ? -1.0 : 1.0;
    c.fz = forc + nl;
    c.lam = scal[SC_LAM * kTNR + nl];
    c.rm = scal[SC_RM * kTNR + nl];
    c.c0 = 0.0;
    c.ld_out = P.ld_out;
    c.ld_in = P.ld_in;
    c.ldo32 = (unsigned)P.ld_out;
    c.ldi32 = (unsigned)P.ld_in;
    c.gout = reinterpret_cast<double*>(P.nlm_out + node0 + nl) + comp;
    c.gin = reinterpret_cast<const double*>(P.nlm_in + node0 + nl) + comp;
    double* wbuf = reinterpret_cast<double*>(bufs) + 2 * nl + comp;
    constexpr size_t kBufD = (size_t)2 * kNRowR * kTNR;                        // doubles per buffer
    c.ap = wbuf + (size_t)(nbuf - 1) * kBufD;

    for (int s = 0; s < P.nstage; ++s) {
        const int ib = s & 1, ob = (s + 1) & 1;
        c.yp = wbuf + (size_t)ib * kBufD;
        c.op = wbuf + (size_t)ob * kBufD;
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
        if (s == 0) {         // computed in front of the symmetry barrier
            c.c0 = -(scal[SC_G0 * kTNR + nl] * scal[SC_C0 * kTNR + nl]);
        } else {
        if (tid < nvalid) {   // <D>(current stage state), one thread per node
            const double2* y = bufs + (size_t)ib * kNRowR * kTNR + tid;
            double2 n2[3], n4[5];
#pragma unroll
            for (int m = 0; m < 3; ++m) n2[m] = y[pslot(2, m) * kTNR];
#pragma unroll
            for (int m = 0; m < 5; ++m) n4[m] = (kL >= 4) ? y[pslot(4, m) * kTNR] : make_double2(0.0, 0.0);
            const double davg = ddrx_davg(global_src(P, node0 + tid), y[0], n2, n4);
            scal[SC_C0 * kTNR + tid] = -(scal[SC_G0 * kTNR + tid] * davg);
        }
        __syncthreads();
        c.c0 = scal[SC_C0 * kTNR + nl];
        }
#endif
        apply_reduced<RIO>(c);
        if (!c.last) __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, SFB_MINB) step_kernel_r(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    step_tile_r<false>(P, smem_raw);
}
// the same step on reduced-form arrays (rows m >= 0 in, rows m >= 0 out)
__global__ void __launch_bounds__(kThreads, SFB_MINB) step_kernel_rr(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    step_tile_r<true>(P, smem_raw);
}

}  // namespace

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    const size_t fixed = (size_t)kNF * kTNR * 16 + (size_t)kNSc * kTNR * 8 + 16;
    const size_t per_buf = (size_t)kNRowR * kTNR * 16;
    const int nbuf_rk = SFB_HORNER ? 2 : 3;
    const size_t smem_max = nbuf_rk * per_buf + fixed;
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
    P.n0_global = 1;
    const int nbuf = P.nstage == 1 ? 1 : nbuf_rk;
    const size_t smem = nbuf * per_buf + fixed;
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
    const long long ntile = (P.N + kTNR - 1) / kTNR;
    if (P.rio) step_kernel_rr<<<(unsigned)ntile, kThreads, smem, st>>>(P);
    else step_kernel_r<<<(unsigned)ntile, kThreads, smem, st>>>(P);
    return cudaGetLastError();
}
