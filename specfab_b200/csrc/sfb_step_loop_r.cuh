// Reduced (one lane per node, real-ODF symmetry) variant of the table-driven loop apply of sfb_step_loop.cuh.
// Included by sfb_step_kernel_r.cuh inside the translation unit's anonymous namespace, after sfb_step_loop.cuh
// (same table, same (mu, chunk) items, same unrolled body).  Differences:
//   * only the plane of rows m >= 0 exists in shared memory (row (l, m) at ((l/2)^2 + m) * kTNR);
//   * a column block with nu = mu - D < 0 is read from the rows (l_j, |nu|) and its partial sums S' are
//     conj-mirrored, S = (-1)^nu conj(S'), by choosing the signs of the four forcing FMAs per item at run time
//     (warp uniform);
//   * rows are finalised for m >= 0 and the mirror rows are written on the last stage.
#pragma once

namespace loopk {

template <bool RIO>
__device__ __forceinline__ void row_out_rt_r(const CtxR& c, int l, int mu, bool rowok, double kr, double ki, double zr, double zi,
                                             double2 n0, double2 acc) {
    double d = fma(c.lam, -(double)(l * (l + 1)), c.c0);
    d = fma(c.rm, c_reg.regdiag[l >> 1], d);
    kr = fma(d, zr, kr);
    ki = fma(d, zi, ki);
    if (mu == 0) { ki = 0.0; zi = 0.0; n0.y = 0.0; acc.y = 0.0; }     // n_l^0 of a real ODF is real
    const int off = ((l >> 1) * (l >> 1) + mu) * kTNR;
    const double n0r = c.first ? zr : n0.x, n0i = c.first ? zi : n0.y;
#if SFB_HORNER
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (rowok && !c.last) c.op[off] = y;
    const double2 res = y;
#else
    const double2 A = make_double2(fma(c.bs, kr, c.first ? zr : acc.x), fma(c.bs, ki, c.first ? zi : acc.y));
    const double2 y = make_double2(fma(c.as, kr, n0r), fma(c.as, ki, n0i));
    if (rowok && !c.last) { c.op[off] = y; c.ap[off] = A; }
    const double2 res = A;
#endif
    if (rowok && c.last && c.valid) {
        if (RIO) {          // reduced-form output array: row (l, mu) at (l/2)^2 + mu, no mirror rows
            c.gout[(long long)((l >> 1) * (l >> 1) + mu) * c.ld_out] = res;
        } else {
            const long long h = (long long)(l * (l + 1) / 2);
            c.gout[(h + mu) * c.ld_out] = res;
            if (mu != 0) c.gout[(h - mu) * c.ld_out] = (mu & 1) ? make_double2(-res.x, res.y) : make_double2(res.x, -res.y);
        }
    }
}

// one forcing item with run-time signs:  ar += fx*sr + gy*si ;  ai += hx*si + fy*sr
//   direct block   (nu >= 0):  fx = f.x, gy = -f.y, hx =  f.x, fy = f.y
//   mirrored block (nu <  0):  fx = s f.x, gy = s f.y, hx = -s f.x, fy = s f.y,  s = (-1)^nu
template <int HBI>
__device__ __forceinline__ void item_r(const double* tp, const double2 f, bool mir, double sg, const double2 (&y)[kNY],
                                       double (&ar)[kCH], double (&ai)[kCH]) {
    const double fx = sg * f.x, fy = sg * f.y;
    const double gy = mir ? fy : -fy, hx = mir ? -fx : fx;
#pragma unroll
    for (int q = 0; q < kCH; ++q) {
        double sr = 0.0, si = 0.0;
#pragma unroll
        for (int b = -HBI; b <= HBI; ++b) {
            const double cf = tp[q * (2 * HBI + 1) + b + HBI];
            sr = fma(cf, y[q + kHB + b].x, sr);
            si = fma(cf, y[q + kHB + b].y, si);
        }
        ar[q] = fma(fx, sr, ar[q]); ar[q] = fma(gy, si, ar[q]);
        ai[q] = fma(hx, si, ai[q]); ai[q] = fma(fy, sr, ai[q]);
    }
}

// fl: the eight M_LROT forcing entries (qe[-2..2], i*qo[-1..1]) of this lane's node, loaded once per stage and kept in
// registers (they are needed by every (mu, chunk) item; the 15 DDRX weights stay in shared memory)
template <int D>
__device__ __forceinline__ void delta_body_r(const CtxR& c, const double2 (&fl)[8], const double2* tp2, int mu, const int (&hh)[kNY],
                                             double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    constexpr int aD = D < 0 ? -D : D;
    constexpr int cnt = body_count<D>();
    const int nu = mu - D;
    const int anu = nu < 0 ? -nu : nu;
    if (anu > kL) return;                                       // warp-uniform: no such column
    double cf[cnt];
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) { const double2 p = tp2[i]; cf[2 * i] = p.x; cf[2 * i + 1] = p.y; }
    const double2* col = c.yp + anu * kTNR;
    double2 y[kNY];
#pragma unroll
    for (int cc = 0; cc < kNY; ++cc) y[cc] = col[hh[cc]];
    if (D == 0) {
#pragma unroll
        for (int q = 0; q < kCH; ++q) { zr[q] = y[q + kHB].x; zi[q] = y[q + kHB].y; }
    }
    const bool mir = nu < 0;
    const double sg = (mir && (anu & 1)) ? -1.0 : 1.0;
    const double* t = cf;
    // register-resident forcing pays for the LROT kernels (-5..8 %); with DDRX the registers are worth more (+6..14 % slower)
    if (aD <= 2) { item_r<1>(t, SFB_DDRX ? c.fz[(D + 2) * kTNR] : fl[aD <= 2 ? D + 2 : 0], mir, sg, y, ar, ai); t += 3 * kCH; }           // A: qe[D]
    if (aD <= 1) { item_r<0>(t, SFB_DDRX ? c.fz[(5 + D + 1) * kTNR] : fl[aD <= 1 ? 5 + D + 1 : 0], mir, sg, y, ar, ai); t += kCH; }       // B: i*qo[D]
#if SFB_DDRX
    if (D == 0) { item_r<0>(t, c.fz[8 * kTNR], mir, sg, y, ar, ai); t += kCH; }                          // lk = 0
    if (aD <= 2) { item_r<1>(t, c.fz[(8 + 3 + D) * kTNR], mir, sg, y, ar, ai); t += 3 * kCH; }           // lk = 2: k = 3 + D
    item_r<2>(t, c.fz[(8 + 10 + D) * kTNR], mir, sg, y, ar, ai);                                          // lk = 4: k = 10 + D
#endif
}

template <int D>
__device__ __forceinline__ void delta_sweep_r(const CtxR& c, const double2 (&fl)[8], const double2* tp2, int mu, const int (&hh)[kNY],
                                              double (&ar)[kCH], double (&ai)[kCH], double (&zr)[kCH], double (&zi)[kCH]) {
    if constexpr (D <= kDm) {
        delta_body_r<D>(c, fl, tp2, mu, hh, ar, ai, zr, zi);
        delta_sweep_r<D + 1>(c, fl, tp2 + body_count<D>() / 2, mu, hh, ar, ai, zr, zi);
    }
}

// ring: this warp's private [2][kPairs] double2 area
template <bool RIO>
__device__ __forceinline__ void apply_loop_r(const CtxR& c, int role, int nroles, double2* ring, int lane) {
    int slot = 0;
    if (role < SFB_LT_NITEMS) ring_fetch(ring, c.ktab + (size_t)role * kPairs, lane);
    double2 fl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) fl[i] = SFB_DDRX ? make_double2(0.0, 0.0) : c.fz[i * kTNR];
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
            hh[cc] = h * h * kTNR;
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
            if (rowok && c.ld_n0) n0[q] = c.gin[(long long)(RIO ? (l >> 1) * (l >> 1) + mu : l * (l + 1) / 2 + mu) * c.ld_in];
#if !SFB_HORNER
            if (rowok && c.ld_acc) acc[q] = c.ap[((l >> 1) * (l >> 1) + mu) * kTNR];
#endif
        }
        delta_sweep_r<-kDm>(c, fl, tp2, mu, hh, ar, ai, zr, zi);
#pragma unroll
        for (int q = 0; q < kCH; ++q) {
            const int l = kL - 2 * (k * kCH + q);
            const bool rowok = l >= mu && l >= 0;
            row_out_rt_r<RIO>(c, rowok ? l : 0, mu, rowok, ar[q], ai[q], zr[q], zi[q], n0[q], acc[q]);
        }
        __syncwarp();          // every lane is done with this slot before it is refilled two items later
        slot ^= 1;
    }
}

}  // namespace loopk
