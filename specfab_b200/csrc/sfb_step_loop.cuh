// Table-driven LOOP form of the operator apply (two lanes per node), included by sfb_step_kernel.cuh when
// SFB_LOOP is defined.  Same (m_i, D)-blocked algorithm as the generated straight-line code
// (codegen/emit_step.py), but the canonical-m sweep and the row chunks are run-time loops over ONE small
// unrolled body, so the code (a few thousand instructions) stays in the instruction cache for any L;
// operator entries stream from a global table (uniform-address 128-bit loads, L1/L2 resident) laid out in
// consumption order by codegen/emit_step.py::emit_loop_table.
//
// Needs: kL, kTN, SFB_DDRX, SFB_CH (rows per chunk), SFB_LT_PER_CHUNK, sfb_lt_chunk_base[], sfb_lt_nchunk[].
#pragma once

namespace loopk {

constexpr int kCH = SFB_CH;
constexpr int kHB = SFB_DDRX ? 2 : 1;               // half band width in rows (|l_i - l_j| <= 2*kHB)
constexpr int kDm = SFB_DDRX ? 4 : 2;
constexpr int kNY = kCH + 2 * kHB;                   // columns held per (chunk, D)

// finalize a row with run-time (l, mu): same arithmetic as row_out<l,mu>
__device__ __forceinline__ void row_out_rt(const Ctx& c, int l, int mu, bool rowok, double kr, double ki, double zr, double zi,
                                           double2 n0, double2 acc) {
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l >> 1], d);
    kr = fma(d, zr, kr);
    ki = fma(d, zi, ki);
    const int off = 2 * ((l >> 1) * (l >> 1) + mu) * kTN;
    const bool own = rowok && ((mu != 0) || c.isA);
    const double n0r = c.first ? zr : n0.x, n0i = c.first ? zi : n0.y;
    const long long goff = (long long)(l * (l + 1) / 2) * c.ld_out + (long long)mu * c.sld;
    double2* outp = (mu == 0 ? c.oz : c.op);
#if SFB_HORNER
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (own && !c.last) outp[off] = y;
    if (own && c.last && c.valid) c.gout[goff] = y;
#else
    double2* accp = (mu == 0 ? c.az : c.ap);
    const double2 A = make_double2(fma(c.bs, kr, c.first ? zr : acc.x), fma(c.bs, ki, c.first ? zi : acc.y));
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (own && !c.last) { outp[off] = y; accp[off] = A; }
    if (own && c.last && c.valid) c.gout[goff] = A;
#endif
}

// one forcing item: rows q = 0..kCH-1, band -HBI..HBI, entries consumed from tp in (q, b) order
template <int HBI>
__device__ __forceinline__ void item(const double* __restrict__& tp, const double2 f, const double2 (&y)[kNY],
                                     double (&ar)[kCH], double (&ai)[kCH]) {
#pragma unroll
    for (int q = 0; q < kCH; ++q) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int b = -HBI; b <= HBI; ++b) {
            const double cf = __ldg(tp + (q * (2 * HBI + 1) + b + HBI));
            sr = fma(cf, y[q + kHB + b].x, sr);
            si = fma(cf, y[q + kHB + b].y, si);
        }
        ar[q] = fma(f.x, sr, ar[q]); ar[q] = fma(-f.y, si, ar[q]);
        ai[q] = fma(f.x, si, ai[q]); ai[q] = fma(f.y, sr, ai[q]);
    }
    tp += kCH * (2 * HBI + 1);
}

template <int D>
__device__ __forceinline__ void delta_body(const Ctx& c, const double* __restrict__& tp, int mu, int k,
                                           double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    constexpr int aD = D < 0 ? -D : D;
    constexpr int cnt = kCH * ((aD <= 2 ? 3 : 0) + (aD <= 1 ? 1 : 0) + (SFB_DDRX ? ((D == 0 ? 1 : 0) + (aD <= 2 ? 3 : 0) + 5) : 0));
    const int nu = mu - D;
    const int anu = nu < 0 ? -nu : nu;
    if (anu > kL) { tp += cnt; return; }                       // warp-uniform: no such column
    const double2* __restrict__ col = (nu > 0) ? c.yp : (nu < 0 ? c.yn : c.yz);
    double2 y[kNY];
#pragma unroll
    for (int cc = 0; cc < kNY; ++cc) {
        const int ci = k * kCH + cc - kHB;                        // top-down column index: l_j = L - 2*ci
        const int lj = kL - 2 * ci;
        const bool ok = (ci >= 0) && (lj >= anu);
        const int h = lj >> 1;
        y[cc] = ok ? col[2 * (h * h + anu) * kTN] : make_double2(0.0, 0.0);
    }
    if (D == 0) {
#pragma unroll
        for (int q = 0; q < kCH; ++q) { zr[q] = y[q + kHB].x; zi[q] = y[q + kHB].y; }
    }
    if (aD <= 2) item<1>(tp, c.fz[(D + 2) * kTN], y, ar, ai);                 // A: qe[D]
    if (aD <= 1) item<0>(tp, c.fz[(5 + D + 1) * kTN], y, ar, ai);             // B: i*qo[D]
#if SFB_DDRX
    if (D == 0) item<0>(tp, c.fz[8 * kTN], y, ar, ai);                        // lk = 0
    if (aD <= 2) item<1>(tp, c.fz[(8 + 3 + D) * kTN], y, ar, ai);             // lk = 2: k = 3 + D
    item<2>(tp, c.fz[(8 + 10 + D) * kTN], y, ar, ai);                         // lk = 4: k = 10 + D
#endif
}

template <int D>
__device__ __forceinline__ void delta_sweep(const Ctx& c, const double* __restrict__& tp, int mu, int k,
                                            double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    if constexpr (D <= kDm) {
        delta_body<D>(c, tp, mu, k, ar, ai, zr, zi);
        delta_sweep<D + 1>(c, tp, mu, k, ar, ai, zr, zi);
    }
}

__device__ __forceinline__ void apply_loop(const Ctx& c) {
    const double* __restrict__ tab = reinterpret_cast<const double*>(c.ktab);
    for (int mu = 0; mu <= kL; ++mu) {
        const int nck = sfb_lt_nchunk[mu];
        const double* __restrict__ tp = tab + (size_t)sfb_lt_chunk_base[mu] * SFB_LT_PER_CHUNK;
        for (int k = 0; k < nck; ++k) {
            double ar[kCH], ai[kCH], zr[kCH], zi[kCH];
            double2 n0[kCH], acc[kCH];
#pragma unroll
            for (int q = 0; q < kCH; ++q) {
                ar[q] = ai[q] = zr[q] = zi[q] = 0.0;
                const int l = kL - 2 * (k * kCH + q);
                const bool rowok = l >= mu && l >= 0;
                n0[q] = make_double2(0.0, 0.0);
                acc[q] = make_double2(0.0, 0.0);
                if (rowok && c.ld_n0 && (mu != 0 || c.isA)) n0[q] = c.gin[(long long)(l * (l + 1) / 2) * c.ld_in + (long long)mu * c.sld_in];
#if !SFB_HORNER
                if (rowok && c.ld_acc) acc[q] = (mu == 0 ? c.az : c.ap)[2 * ((l >> 1) * (l >> 1) + mu) * kTN];
#endif
            }
            delta_sweep<-kDm>(c, tp, mu, k, ar, ai, zr, zi);
#pragma unroll
            for (int q = 0; q < kCH; ++q) {
                const int l = kL - 2 * (k * kCH + q);
                const bool rowok = l >= mu && l >= 0;
                row_out_rt(c, rowok ? l : 0, mu, rowok, ar[q], ai[q], zr[q], zi[q], n0[q], acc[q]);
            }
        }
    }
}

}  // namespace loopk
