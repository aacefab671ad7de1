// Table-driven LOOP form of the operator apply (two lanes per node), included by sfb_step_kernel.cuh when
// SFB_LOOP is defined.  Same (m_i, D)-blocked algorithm as the generated straight-line code
// (codegen/emit_step.py), but the canonical-m sweep and the row chunks are run-time loops over ONE small
// unrolled body, so the code (a few thousand instructions) stays in the instruction cache for any L;
// operator entries stream from a global table (uniform-address 128-bit loads, L1/L2 resident) laid out in
// consumption order by codegen/emit_step.py::emit_loop_table.
//
// The work items (mu, chunk) all cost the same and are dealt round-robin to the SFB_R warp roles that share a
// node group, so the number of warps per resident node is a free tuning parameter (the code is shared).
// Column loads are unpredicated: a column that does not exist is read from a clamped, always-initialised
// address and multiplied by a zero table entry (the unused m=0 twin rows are zero-filled by the skeleton).
//
// Needs: kL, kTN, SFB_DDRX, SFB_CH (rows per chunk), SFB_LT_PER_CHUNK, SFB_LT_NITEMS, sfb_lt_item_mu[], sfb_lt_item_k[].
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
__device__ __forceinline__ void item(const double* tp, const double2 f, const double2 (&y)[kNY],
                                     double (&ar)[kCH], double (&ai)[kCH]) {
#pragma unroll
    for (int q = 0; q < kCH; ++q) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int b = -HBI; b <= HBI; ++b) {
            const double cf = tp[q * (2 * HBI + 1) + b + HBI];
            sr = fma(cf, y[q + kHB + b].x, sr);
            si = fma(cf, y[q + kHB + b].y, si);
        }
        ar[q] = fma(f.x, sr, ar[q]); ar[q] = fma(-f.y, si, ar[q]);
        ai[q] = fma(f.x, si, ai[q]); ai[q] = fma(f.y, sr, ai[q]);
    }
}

// The table entries of an item are streamed by the warp itself into a private double-buffered ring in shared
// memory with cp.async (16 B per lane) one item ahead of their use, then read as broadcast LDS.128.
template <int D>
__host__ __device__ constexpr int body_count() {
    constexpr int aD = D < 0 ? -D : D;
    return kCH * ((aD <= 2 ? 3 : 0) + (aD <= 1 ? 1 : 0) + (SFB_DDRX ? ((D == 0 ? 1 : 0) + (aD <= 2 ? 3 : 0) + 5) : 0));
}

template <int D>
__device__ __forceinline__ void delta_body(const Ctx& c, const double2* tp2, int mu, const int (&hh)[kNY],
                                           double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    constexpr int aD = D < 0 ? -D : D;
    constexpr int cnt = body_count<D>();
    static_assert(cnt % 2 == 0, "entries per D body must be even");
    const int nu = mu - D;
    const int anu = nu < 0 ? -nu : nu;
    if (anu > kL) return;                                       // warp-uniform: no such column
    double cf[cnt];
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) { const double2 p = tp2[i]; cf[2 * i] = p.x; cf[2 * i + 1] = p.y; }
    const double2* col = ((nu > 0) ? c.yp : (nu < 0 ? c.yn : c.yz)) + anu * (2 * kTN);
    double2 y[kNY];
#pragma unroll
    for (int cc = 0; cc < kNY; ++cc) y[cc] = col[hh[cc]];
    if (D == 0) {
#pragma unroll
        for (int q = 0; q < kCH; ++q) { zr[q] = y[q + kHB].x; zi[q] = y[q + kHB].y; }
    }
    const double* t = cf;
    if (aD <= 2) { item<1>(t, c.fz[(D + 2) * kTN], y, ar, ai); t += 3 * kCH; }                 // A: qe[D]
    if (aD <= 1) { item<0>(t, c.fz[(5 + D + 1) * kTN], y, ar, ai); t += kCH; }               // B: i*qo[D]
#if SFB_DDRX
    if (D == 0) { item<0>(t, c.fz[8 * kTN], y, ar, ai); t += kCH; }                          // lk = 0
    if (aD <= 2) { item<1>(t, c.fz[(8 + 3 + D) * kTN], y, ar, ai); t += 3 * kCH; }           // lk = 2: k = 3 + D
    item<2>(t, c.fz[(8 + 10 + D) * kTN], y, ar, ai);                                          // lk = 4: k = 10 + D
#endif
}

template <int D>
__device__ __forceinline__ void delta_sweep(const Ctx& c, const double2* tp2, int mu, const int (&hh)[kNY],
                                            double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    if constexpr (D <= kDm) {
        delta_body<D>(c, tp2, mu, hh, ar, ai, zr, zi);
        delta_sweep<D + 1>(c, tp2 + body_count<D>() / 2, mu, hh, ar, ai, zr, zi);
    }
}

constexpr int kPairs = SFB_LT_PER_CHUNK / 2;             // double2 entries per item

__device__ __forceinline__ void ring_fetch(double2* ring, const double2* __restrict__ src, int lane) {
#pragma unroll
    for (int i = 0; i < (kPairs + 31) / 32; ++i) {
        const int e = lane + 32 * i;
        if (e < kPairs) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + e) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ring: this warp's private [2][kPairs] double2 area
__device__ __forceinline__ void apply_loop(const Ctx& c, int role, int nroles, double2* ring, int lane) {
    int slot = 0;
    if (role < SFB_LT_NITEMS) ring_fetch(ring, c.ktab + (size_t)role * kPairs, lane);
    for (int it = role; it < SFB_LT_NITEMS; it += nroles) {
        const int nxt = it + nroles;
        if (nxt < SFB_LT_NITEMS) {
            ring_fetch(ring + (slot ^ 1) * kPairs, c.ktab + (size_t)nxt * kPairs, lane);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const double2* tp2 = ring + slot * kPairs;
        const int mu = sfb_lt_item_mu[it], k = sfb_lt_item_k[it];
        int hh[kNY];
#pragma unroll
        for (int cc = 0; cc < kNY; ++cc) {
            int h = kL / 2 - (k * kCH + cc - kHB);                   // l_j / 2 of the column (top-down index)
            h = h > kL / 2 ? kL / 2 : (h < 0 ? 0 : h);               // clamp: non-existent columns meet zero entries
            hh[cc] = 2 * h * h * kTN;
        }
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
            if (rowok && c.ld_acc && (mu != 0 || c.isA)) acc[q] = (mu == 0 ? c.az : c.ap)[2 * ((l >> 1) * (l >> 1) + mu) * kTN];   // set B never owns m = 0 rows
#endif
        }
        delta_sweep<-kDm>(c, tp2, mu, hh, ar, ai, zr, zi);
#pragma unroll
        for (int q = 0; q < kCH; ++q) {
            const int l = kL - 2 * (k * kCH + q);
            const bool rowok = l >= mu && l >= 0;
            row_out_rt(c, rowok ? l : 0, mu, rowok, ar[q], ai[q], zr[q], zi[q], n0[q], acc[q]);
        }
        __syncwarp();          // every lane is done with this slot before it is refilled two items later
        slot ^= 1;
    }
}

}  // namespace loopk
