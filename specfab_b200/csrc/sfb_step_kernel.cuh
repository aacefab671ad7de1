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
constexpr int kNSc = 20;                                  // per-node scalars
static_assert(kTN % 16 == 0, "tile must be a multiple of 16 nodes");

__constant__ SfbRegConst c_reg;

__host__ __device__ constexpr int pslot(int l, int a) { return (l / 2) * (l / 2) + a; }
__host__ __device__ constexpr int hrow(int l) { return l * (l + 1) / 2; }

// scalar slots
enum { SC_C0 = 0, SC_LAM = 1, SC_RM = 2, SC_G0 = 3, SC_TAUV = 4, SC_TSQV = 10, SC_NORM = 16 };

struct Ctx {
    const double2 *yz, *yp, *yn;      // stage input planes (zero / positive / negative canonical m)
    const double2* fz;                // forcing block of this lane set
    double2 *oz, *op;                 // next-stage buffer (rows owned by this lane: zero / positive plane)
    const double2 *nz, *np_;          // n0 buffer
    double2 *az, *ap;                 // RK accumulator buffer
    double2* gout;                    // global output, already offset by node
    const double2* gin;               // global input (n0 re-read when the 4th smem buffer does not fit)
    long long ld_out, sld;            // row stride, signed row stride (+ld for set A, -ld for set B)
    long long ld_in, sld_in;
    double c0, lam, rm;               // diagonal: c0 + lam*(-l(l+1)) + rm*regdiag_l
    double as, bs;                    // stage coefficients: ynext = n0 + as*k ; acc += bs*k
    bool first, last, isA, valid, n0g;
};

template <int l, int mu>
__device__ __forceinline__ void row_out(const Ctx& c, double kr, double ki, double zr, double zi) {
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l / 2], d);
    kr = fma(d, zr, kr);
    ki = fma(d, zi, ki);
    constexpr int off = 2 * pslot(l, mu) * kTN;
    if (mu == 0 && !c.isA) return;   // m = 0 rows are computed by both lane sets; set A owns them
    if (c.first) {                   // n0 is the stage input itself
        const double2 A = make_double2(fma(c.bs, kr, zr), fma(c.bs, ki, zi));
        if (c.last) {                // Euler
            if (c.valid) c.gout[(long long)hrow(l) * c.ld_out + (long long)mu * c.sld] = A;
        } else {
            (mu == 0 ? c.az : c.ap)[off] = A;
            (mu == 0 ? c.oz : c.op)[off] = make_double2(fma(c.as, kr, zr), fma(c.as, ki, zi));
        }
    } else {
        double2 A = (mu == 0 ? c.az : c.ap)[off];
        A.x = fma(c.bs, kr, A.x);
        A.y = fma(c.bs, ki, A.y);
        if (c.last) {
            if (c.valid) c.gout[(long long)hrow(l) * c.ld_out + (long long)mu * c.sld] = A;
        } else {
            (mu == 0 ? c.az : c.ap)[off] = A;
            double2 n0;
            if (c.n0g) n0 = c.valid ? c.gin[(long long)hrow(l) * c.ld_in + (long long)mu * c.sld_in] : make_double2(0.0, 0.0);
            else n0 = (mu == 0 ? c.nz : c.np_)[off];
            (mu == 0 ? c.oz : c.op)[off] = make_double2(fma(c.as, kr, n0.x), fma(c.as, ki, n0.y));
        }
    }
}
#define SFB_ROW_OUT(l, mu, ar, ai, zr, zi) row_out<l, mu>(c, ar, ai, zr, zi)
// keeps the warps of a CTA within one instruction-cache window of the generated straight-line code
#define SFB_LOCKSTEP() __syncthreads()

__device__ __forceinline__ void apply_role(const Ctx& c, int role) {
    const double2* __restrict__ yz = c.yz;
    const double2* __restrict__ yp = c.yp;
    const double2* __restrict__ yn = c.yn;
    const double2* __restrict__ fz = c.fz;
#include SFB_APPLY_INC
}

// ---- complex helpers (forcing preparation) ----
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 cneg(double2 a) { return make_double2(-a.x, -a.y); }

// src/dynamics.f90:563-579 ; q[0..4] <-> m = -2..2 ; M symmetric, row-major m[3][3]
__device__ __forceinline__ void quad_rr(const double m[3][3], double2 q[5]) {
    const double fsq = 0x1.4b5eee37a973cp-1;    // sqrt(2*Pi/15)
    const double sp5 = 0x1.95d83f429fefap-1;    // sqrt(Pi/5)
    const double xx = m[0][0], yy = m[1][1], zz = m[2][2], xy = m[0][1], xz = m[0][2], yz = m[1][2];
    q[0] = make_double2(fsq * (xx - yy), fsq * (2 * xy));
    q[1] = make_double2((2 * fsq) * xz, (2 * fsq) * yz);
    q[2] = make_double2(-((SFB_TWOTHIRDS_F * sp5) * (xx + yy - 2 * zz)), 0.0);
    q[3] = make_double2(-((2 * fsq) * xz), (2 * fsq) * yz);
    q[4] = make_double2(fsq * (xx - yy), fsq * (-2 * xy));
}

// src/dynamics.f90:581-593 ; q[0..2] <-> m = -1..1 ; M antisymmetric
__device__ __forceinline__ void quad_tp(const double m[3][3], double2 q[3]) {
    const double fsq1 = 0x1.727bdd17583bbp+0;   // sqrt(2*Pi/3)
    const double xy = m[0][1], xz = m[0][2], yz = m[1][2];
    q[0] = make_double2(fsq1 * yz, fsq1 * (-xz));
    q[1] = make_double2(fsq1 * (SFB_SQRT2_F * xy), 0.0);
    q[2] = make_double2(fsq1 * (-yz), fsq1 * (-xz));
}

#if SFB_DDRX
// src/include/ddrx-coupling-weights.f90:1-16 with real(4) constants; qt**(2.0) == qt*qt
__device__ __forceinline__ void ddrx_weights_raw(const double2 qt[5], double2 g[15]) {
    const double s5 = 0x1.1e377ap+1, s15 = 0x1.3988e2p+0, s6 = 0x1.3988e2p+1, s2 = 0x1.6a09e6p+0, s3 = 0x1.bb67aep+0;
    const double c2s14 = 0x1.deeea2p+2;   // 2*Sqrt((14.0))   real(4)
    const double c4s7 = 0x1.52a7fap+3;    // 4*Sqrt((7.0))    real(4)
    const double c3s5 = 0x1.ad5338p+2;    // 3.*Sqrt((5.0))   real(4)
    const double2 qm2 = qt[0], qm1 = qt[1], q0 = qt[2], qp1 = qt[3], qp2 = qt[4];
    const double2 q0q0 = cmul(q0, q0), qm1qm1 = cmul(qm1, qm1), qp1qp1 = cmul(qp1, qp1);
    double2 t;
    t = cadd(cadd(q0q0, cmul(cscale(-2.0, qm1), qp1)), cmul(cscale(2.0, qm2), qp2));
    t = cscale(7.0, t); g[0] = make_double2(t.x / s5, t.y / s5);
    g[1] = cadd(cscale(s15, qm1qm1), cmul(cscale(-2.0, q0), qm2));
    g[2] = cadd(cmul(q0, qm1), cmul(cscale(-s6, qp1), qm2));
    g[3] = cadd(cadd(q0q0, cmul(cscale(-1.0, qm1), qp1)), cmul(cscale(-2.0, qm2), qp2));
    g[4] = cadd(cmul(q0, qp1), cmul(cscale(-s6, qm1), qp2));
    g[5] = cadd(cscale(s15, qp1qp1), cmul(cscale(-2.0, q0), qp2));
    t = cneg(cscale(c2s14, cmul(qm2, qm2))); g[6] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cmul(cscale(c4s7, qm1), qm2)); g[7] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cscale(s2, qm1qm1), cmul(cscale(s3, q0), qm2)))); g[8] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cmul(cscale(s6, q0), qm1), cmul(qp1, qm2)))); g[9] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cadd(cscale(3.0, q0q0), cmul(cscale(4.0, qm1), qp1)), cmul(qm2, qp2))));
    g[10] = make_double2(t.x / c3s5, t.y / c3s5);
    t = cneg(cscale(4.0, cadd(cmul(cscale(s6, q0), qp1), cmul(qm1, qp2)))); g[11] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cscale(s2, qp1qp1), cmul(cscale(s3, q0), qp2)))); g[12] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cmul(cscale(c4s7, qp1), qp2)); g[13] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(c2s14, cmul(qp2, qp2))); g[14] = make_double2(t.x / 3.0, t.y / 3.0);
}
#endif

// catalyst index k of (lk,mk) and of its mirror (lk,-mk); k = 0 | 1..5 | 6..14
__device__ __forceinline__ int cat_mirror(int k) { return k == 0 ? 0 : (k < 6 ? 6 - k + 0 : 20 - k); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Per-node forcing preparation: quadric coefficients, DDRX weights, diagonal scalars.
__device__ void prep_node(const SfbStepParams& P, long long node, int t, double2* forc, double* scal) {
    double u[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) u[i][j] = P.ugrad[(long long)(i + 3 * j) * P.ld_u + node];
    double D[3][3], W[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { D[i][j] = (u[i][j] + u[j][i]) / 2; W[i][j] = (u[i][j] - u[j][i]) / 2; }
    double2* fA = forc + t;                    // lane set A block: entry e at fA[e*kTN]
    double2* fB = forc + kNF * kTN + t;        // lane set B block
    {   // ---- M_LROT weights, src/dynamics.f90:71-76
        double sq[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) sq[i][j] = D[i][0] * D[0][j] + D[i][1] * D[1][j] + D[i][2] * D[2][j];
        const double zetanorm = P.zeta / sqrt(sq[0][0] + sq[1][1] + sq[2][2]);   // 0/0 -> NaN like the reference
        double E[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) E[i][j] = P.iota * D[i][j] + zetanorm * sq[i][j];
        double2 qe[5], qo[3];
        quad_rr(E, qe);
        quad_tp(W, qo);
        if (!P.use_lrot) {
#pragma unroll
            for (int i = 0; i < 5; ++i) qe[i] = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 3; ++i) qo[i] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int d = -2; d <= 2; ++d) { fA[(d + 2) * kTN] = qe[d + 2]; fB[(d + 2) * kTN] = qe[-d + 2]; }
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
            const double2 wa = make_double2(-qo[d + 1].y, qo[d + 1].x);      //  i*qo[d]
            const double2 wb = make_double2(qo[-d + 1].y, -qo[-d + 1].x);    // -i*qo[-d]
            fA[(5 + d + 1) * kTN] = wa;
            fB[(5 + d + 1) * kTN] = wb;
        }
        // M_REG: -nu*||D||_F * regdiag   src/dynamics.f90:516-517
        double fro = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) fro += D[i][j] * D[i][j];
        scal[SC_RM * kTN + t] = P.use_reg ? -(P.nu_mult * (c_reg.nu * sqrt(fro))) : 0.0;
    }
    scal[SC_LAM * kTN + t] = P.lambda_arr ? P.lambda_arr[node] : P.lambda;
    scal[SC_C0 * kTN + t] = 0.0;
#if SFB_DDRX
    {
        const double g0 = P.gamma0_arr ? P.gamma0_arr[node] : P.gamma0;
        scal[SC_G0 * kTN + t] = g0;
        double T[3][3];
        if (P.tau) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) T[i][j] = P.tau[(long long)(i + 3 * j) * P.ld_t + node];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) T[i][j] = D[i][j];
        }
        double2 qt[5], g[15];
        quad_rr(T, qt);
        ddrx_weights_raw(qt, g);
        double dd = 0.0;                                   // doubleinner22(tau,tau) = tau_ij tau_ji
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) dd += T[i][j] * T[j][i];
        const double kk = 0x1.14d2dcd9ceb17p-3;            // (3*Sqrt(5/Pi))/28.
#pragma unroll
        for (int k = 0; k < 15; ++k) {                     // g = k*g*5/(tau:tau)  (src/dynamics.f90:293), then *Gamma0
            double2 v = make_double2(((kk * g[k].x) * 5) / dd, ((kk * g[k].y) * 5) / dd);
            g[k] = make_double2(g0 * v.x, g0 * v.y);
        }
#pragma unroll
        for (int k = 0; k < 15; ++k) { fA[(8 + k) * kTN] = g[k]; fB[(8 + k) * kTN] = g[cat_mirror(k)]; }
        // <D> ingredients, src/dynamics.f90:415-417
        double sq[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) sq[i][j] = T[i][0] * T[0][j] + T[i][1] * T[1][j] + T[i][2] * T[2][j];
        const double tv[6] = {T[0][0], T[1][1], T[2][2], SFB_SQRT2 * T[1][2], SFB_SQRT2 * T[0][2], SFB_SQRT2 * T[0][1]};
        const double sv[6] = {sq[0][0], sq[1][1], sq[2][2], SFB_SQRT2 * sq[1][2], SFB_SQRT2 * sq[0][2], SFB_SQRT2 * sq[0][1]};
#pragma unroll
        for (int p = 0; p < 6; ++p) { scal[(SC_TAUV + p) * kTN + t] = tv[p]; scal[(SC_TSQV + p) * kTN + t] = sv[p]; }
        scal[SC_NORM * kTN + t] = sq[0][0] + sq[1][1] + sq[2][2];
    }
#endif
}

__global__ void __launch_bounds__(kThreads, SFB_MINB) step_kernel(const SfbStepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nbuf = P.nstage == 1 ? 1 : (P.n0_global ? 3 : 4);
    double2* bufs = reinterpret_cast<double2*>(smem_raw);
    double2* forc = bufs + (size_t)nbuf * kNRow * kTN;
    double* scal = reinterpret_cast<double*>(forc + 2 * kNF * kTN);
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(scal + kNSc * kTN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = warp / kR, role = warp % kR;
    const int sb = lane >> 4;                       // 0: lane set A (m>=0), 1: lane set B (m<=0)
    const int nl = group * 16 + (lane & 15);        // node within tile
    const long long node0 = (long long)blockIdx.x * kTN;
    const int nvalid = (int)min((long long)kTN, P.N - node0);

    // ---- stage the tile: one bulk copy per coefficient row into buffer 0
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
    // ---- meanwhile: per-node forcing
    if (tid < nvalid) prep_node(P, node0 + tid, tid, forc, scal);
    {   // wait for the tile
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mb) : "memory");
        }
    }
    __syncthreads();

    Ctx c;
    c.isA = (sb == 0);
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
    c.n0g = P.n0_global != 0;
    double2* b0 = bufs + nl;
    c.nz = b0; c.np_ = b0 + sb * kTN;
    double2* acc = bufs + (size_t)(nbuf - 1) * kNRow * kTN + nl;
    c.az = acc; c.ap = acc + sb * kTN;

    for (int s = 0; s < P.nstage; ++s) {
        // 4 buffers (0 = n0, 3 = acc): s=0: 0 -> 1 ; s=1: 1 -> 2 ; s=2: 2 -> 1 ; s=3: in 1
        // 3 buffers (n0 re-read from global, 2 = acc): inputs 0,1,0,1 ; outputs 1,0,1
        const int ib = P.n0_global ? (s & 1) : ((s == 0) ? 0 : (s == 2 ? 2 : 1));
        const int ob = P.n0_global ? ((s + 1) & 1) : ((s == 1) ? 2 : 1);
        const double2* yin = bufs + (size_t)ib * kNRow * kTN + nl;
        double2* yout = bufs + (size_t)ob * kNRow * kTN + nl;
        c.yz = yin; c.yp = yin + sb * kTN; c.yn = yin + (1 - sb) * kTN;
        c.oz = yout; c.op = yout + sb * kTN;
        c.first = (s == 0);
        c.last = (s == P.nstage - 1);
        if (P.nstage == 1) { c.as = 0.0; c.bs = P.dt; }
        else {
            c.as = (s == 2) ? P.dt : 0.5 * P.dt;
            c.bs = (s == 0 || s == 3) ? P.dt / 6 : P.dt / 3;
        }
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
        apply_role(c, role);
        if (!c.last) __syncthreads();
    }
}

}  // namespace

extern "C" cudaError_t SFB_NAME(const SfbStepParams& Pin, const SfbRegConst& reg, cudaStream_t st) {
    static bool attr_done[64] = {false};
    const size_t fixed = (size_t)2 * kNF * kTN * 16 + (size_t)kNSc * kTN * 8 + 16;
    const size_t per_buf = (size_t)kNRow * kTN * 16;
    const size_t lim = 227 * 1024;
    const int nbuf_rk = (4 * per_buf + fixed <= lim) ? 4 : 3;
    const size_t smem_max = (nbuf_rk * per_buf + fixed <= lim) ? nbuf_rk * per_buf + fixed : per_buf + fixed;
    cudaError_t e;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_done[dev]) {
        e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    SfbStepParams P = Pin;
    P.n0_global = (P.nstage != 1 && nbuf_rk == 3) ? 1 : 0;
    const int nbuf = P.nstage == 1 ? 1 : nbuf_rk;
    const size_t smem = nbuf * per_buf + fixed;
    if (smem > lim) return cudaErrorInvalidConfiguration;
    {   // regularisation constants live in this unit's __constant__ bank; refresh when they change
        static SfbRegConst last[64];
        static bool have[64] = {false};
        if (!have[dev] || memcmp(&last[dev], &reg, sizeof(SfbRegConst)) != 0) {
            e = cudaMemcpyToSymbolAsync(c_reg, &reg, sizeof(SfbRegConst), 0, cudaMemcpyHostToDevice, st);
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
