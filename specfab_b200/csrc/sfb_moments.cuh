// Structure-tensor moments of the spectral state: device functions shared by the step kernel
// (<D> of DDRX) and the a2/a4/eig/Eij kernels.  Reference: src/moments.f90:37-55,164-218 with
// src/include/ev_c2__body.f90, ev_c4_Mandel__body.f90 (d0 constants) and ev_c4__body.f90
// (real(4) constants).  Arithmetic order follows the oracle restatement (oracle/specfab_oracle.py).
#pragma once
#include "sfb_common.cuh"

namespace sfb {

// complex(8) division the way GCC expands it under -fcx-fortran-rules (Smith's range reduction)
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
    double ratio, div, tr, ti;
    if (fabs(b.x) < fabs(b.y)) {
        ratio = b.x / b.y; div = (b.x * ratio) + b.y;
        tr = (a.x * ratio) + a.y; ti = (a.y * ratio) - a.x;
    } else {
        ratio = b.y / b.x; div = (b.y * ratio) + b.x;
        tr = (a.y * ratio) + a.x; ti = a.y - (a.x * ratio);
    }
    return make_double2(tr / div, ti / div);
}

// Mandel order (11,22,33,sqrt2*23,sqrt2*13,sqrt2*12)   src/mandel.f90:15-24
#define SFB_SQRT2 1.4142135623730951

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// a2 as Mandel 6-vector.  n2: n_2^m for m = 0,1,2 (only m>=0 enters, ev_c2__body.f90:1-17).
// n2mhat = n2m/n00 is formed with ONE complex reciprocal (3 divisions instead of 9; <= 2 ulp from the reference's divisions);
// hout (optional) receives the three quotients.
__device__ __forceinline__ void ev_c2_mandel(double2 n00, double2 n20, double2 n21, double2 n22, double a2v[6], double2* hout = nullptr) {
    const double2 rinv = cdiv(make_double2(1.0, 0.0), n00);
    const double2 h0 = cmul2(n20, rinv), h1 = cmul2(n21, rinv), h2 = cmul2(n22, rinv);
    if (hout) { hout[0] = h0; hout[1] = h1; hout[2] = h2; }
    const double c = 0.5 * 0.816496580927726;      // 0.5d0*sqrt(2.0d0/3)
    const double s215 = 0.3651483716701107;        // sqrt(2/15.0d0)
    const double third = 1.0 / 3.0;
    const double e11 = -(c * h0.x) + h2.x;
    const double e22 = -(c * h0.x) - h2.x;
    const double e33 = 0.816496580927726 * h0.x;
    const double e12 = -h2.y, e13 = -h1.x, e23 = h1.y;
    a2v[0] = s215 * e11 + third;
    a2v[1] = s215 * e22 + third;
    a2v[2] = s215 * e33 + third;
    a2v[3] = SFB_SQRT2 * (s215 * e23);
    a2v[4] = SFB_SQRT2 * (s215 * e13);
    a2v[5] = SFB_SQRT2 * (s215 * e12);
}

// index of (a,b), a<=b, in the packed upper triangle of a symmetric 6x6
__host__ __device__ __forceinline__ constexpr int tri6(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }

// a4 as packed symmetric 6x6 Mandel matrix (21 entries).  ev_c4_Mandel__body.f90:1-37
// n2[m], m=0..2 ; n4[m], m=0..4 (only m>=0 coefficients enter this body).
__device__ __forceinline__ void ev_c4_mandel(double2 n00, const double2 n2[3], const double2 n4[5], double e[21]) {
    const double s5 = 2.23606797749979, s6 = 2.449489742783178, s7 = 2.6457513110645907, s10 = 3.1622776601683795,
                 s30 = 5.477225575051661, s70 = 8.366600265340756, s15 = 3.872983346207417, s3 = 1.7320508075688772;
    const double p315 = 5.196152422706632;      // (3.0d0)**1.5
    const double r00 = n00.x;
    const double r20 = n2[0].x, r21 = n2[1].x, r22 = n2[2].x, i21 = n2[1].y, i22 = n2[2].y;
    const double r40 = n4[0].x, r41 = n4[1].x, r42 = n4[2].x, r43 = n4[3].x, r44 = n4[4].x;
    const double i41 = n4[1].y, i42 = n4[2].y, i43 = n4[3].y, i44 = n4[4].y;
    e[tri6(0, 0)] = 21.0 * r00 + (s5 * -6.0) * r20 + (s30 * 6.0) * r22 + 3.0 * r40 + (s10 * -2.0) * r42 + s70 * r44;
    e[tri6(0, 1)] = 7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44;
    e[tri6(0, 2)] = 7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42;
    e[tri6(0, 3)] = s10 * (s6 * i21 + -1.0 * i41 + s7 * i43);
    e[tri6(0, 4)] = (s10 * -1.0) * ((3.0 * s6) * r21 + -3.0 * r41 + s7 * r43);
    e[tri6(0, 5)] = (-2.0 * s5) * (p315 * i22 + -1.0 * i42 + s7 * i44);
    e[tri6(1, 1)] = 21.0 * r00 + (s5 * -6.0) * (r20 + s6 * r22) + 3.0 * r40 + s10 * (2.0 * r42 + s7 * r44);
    e[tri6(1, 2)] = 7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42);
    e[tri6(1, 3)] = s10 * ((3.0 * s6) * i21 + -3.0 * i41 + (-1.0 * s7) * i43);
    e[tri6(1, 4)] = (s15 * -2.0) * r21 + s10 * (r41 + s7 * r43);
    e[tri6(1, 5)] = (2.0 * s5) * ((-3.0 * s3) * i22 + i42 + s7 * i44);
    e[tri6(2, 2)] = 21.0 * r00 + (12.0 * s5) * r20 + 8.0 * r40;
    e[tri6(2, 3)] = s10 * ((3.0 * s6) * i21 + 4.0 * i41);
    e[tri6(2, 4)] = (s10 * -1.0) * ((3.0 * s6) * r21 + 4.0 * r41);
    e[tri6(2, 5)] = (-2.0 * s5) * (s3 * i22 + 2.0 * i42);
    e[tri6(3, 3)] = 2.0 * (7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42));
    e[tri6(3, 4)] = (s10 * -2.0) * (s3 * i22 + 2.0 * i42);
    e[tri6(3, 5)] = s5 * ((-2.0 * s6) * r21 + 2.0 * (r41 + s7 * r43));
    e[tri6(4, 4)] = 2.0 * (7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42);
    e[tri6(4, 5)] = (2.0 * s5) * (s6 * i21 + -1.0 * i41 + s7 * i43);
    e[tri6(5, 5)] = 2.0 * (7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44);
    const double k = 0x1.149200acaee94p-5;       // (2*Sqrt(Pi))/105. = 0.03376102573153364
    const double c0 = 3.5449077018110318 * r00;  // f_ev_c0 = REAL(sqrt(4*Pi)*n00)  src/moments.f90:184-189
    const double kc = k / c0;                    // ev*k/f_ev_c0: one division instead of 21 (differs by <= 1 ulp)
#pragma unroll
    for (int q = 0; q < 21; ++q) e[q] = e[q] * kc;
}

// <D> = 5[(tau.tau):a2 - tau:a4:tau]/(tau:tau)       src/dynamics.f90:402-422
// tauv, tsqv: Mandel vectors of tau and tau.tau ; norm = tr(tau.tau)
__device__ __forceinline__ double ev_D2(double2 n00, const double2 n2[3], const double2 n4[5],
                                        const double tauv[6], const double tsqv[6], double norm) {
    double a2v[6], e[21];
    ev_c2_mandel(n00, n2[0], n2[1], n2[2], a2v);
    ev_c4_mandel(n00, n2, n4, e);
    double d1 = 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) d1 += tsqv[p] * a2v[p];
    double d2 = 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        double r = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) r += e[p <= q ? tri6(p, q) : tri6(q, p)] * tauv[q];
        d2 += tauv[p] * r;
    }
    return 5 * (d1 - d2) / norm;
}

}  // namespace sfb
