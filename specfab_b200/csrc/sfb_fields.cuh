// Per-node field evaluations on the resident spectral state: a2, a4, eigenframe, Eij.
// One thread per node, node-contiguous (coalesced) loads of the first 15 coefficient rows.
// Reference: src/moments.f90:37-55 (a2, a4), src/frames.f90:14-80 (eig, eigframe, eig3),
// src/enhancementfactors.f90:23-69,398-413 (Eij_tranisotropic, Evw, tau_vv/vw),
// src/homogenizations.f90:69-116,145-258 (Sachs / Taylor, n'=1), src/rheologies.f90:123-135.
#pragma once
#include "sfb_common.cuh"
#include "sfb_moments.cuh"
#include "specfab_b200.h"

namespace sfb {

// ---------------------------------------------------------------------------------------------
// a4 with the reference's real(4) constants (src/include/ev_c4__body.f90:1-94): 15 unique entries
// u[q], q indexing the sorted index quadruples
//  0:1111 1:1112 2:1113 3:1122 4:1123 5:1133 6:1222 7:1223 8:1233 9:1333 10:2222 11:2223 12:2233 13:2333 14:3333
// n2[0..4] <-> m=-2..2, n4[0..8] <-> m=-4..4.  Only REAL(...) of each complex sum is needed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ev_c4_unique(double2 n00, const double2 n2[5], const double2 n4[9], double u[15]) {
    // real(4) constants promoted to double (oracle/specfab_oracle.py f_ev_c4)
    const double s5 = 0x1.1e377ap+1, s30 = 0x1.5e8adep+2, s10 = 0x1.94c584p+1, s70 = 0x1.0bbb3p+3, s2 = 0x1.6a09e6p+0,
                 s3 = 0x1.bb67aep+0, s7 = 0x1.52a7fap+1, s6 = 0x1.3988e2p+1;
    const double m12s5 = -0x1.ad5338p+4, s30x6 = 0x1.06e826p+5, s10x2 = 0x1.94c584p+2;
    const double m3s3 = -0x1.4c8dc2p+2, p3_15 = 0x1.4c8dc2p+2;
    const double p3s6 = 0x1.d64d54p+2, m4s5 = -0x1.1e377ap+3, p2s5 = 0x1.1e377ap+2, p12s5 = 0x1.ad5338p+4;
#define RE2(m) n2[(m) + 2].x
#define IM2(m) n2[(m) + 2].y
#define RE4(m) n4[(m) + 4].x
#define IM4(m) n4[(m) + 4].y
    // entries of the form REAL(Z / s5)
    u[0] = (42.0 * n00.x + m12s5 * RE2(0) + s30x6 * RE2(-2) + s30x6 * RE2(2) + 6.0 * RE4(0) + (-s10x2) * RE4(-2) + (-s10x2) * RE4(2) + s70 * RE4(-4) + s70 * RE4(4)) / s5;
    u[3] = (14.0 * n00.x + m4s5 * RE2(0) + 2.0 * RE4(0) + (-s70) * RE4(-4) + (-s70) * RE4(4)) / s5;
    u[5] = (14.0 * n00.x + p2s5 * RE2(0) + s30 * RE2(-2) + s30 * RE2(2) + -8.0 * RE4(0) + s10x2 * RE4(-2) + s10x2 * RE4(2)) / s5;
    u[10] = (42.0 * n00.x + m12s5 * RE2(0) + (-s30x6) * RE2(-2) + (-s30x6) * RE2(2) + 6.0 * RE4(0) + s10x2 * RE4(-2) + s10x2 * RE4(2) + s70 * RE4(-4) + s70 * RE4(4)) / s5;
    u[12] = (14.0 * n00.x + p2s5 * RE2(0) + (-s30) * RE2(-2) + (-s30) * RE2(2) + -8.0 * RE4(0) + (-s10x2) * RE4(-2) + (-s10x2) * RE4(2)) / s5;
    u[14] = (2.0 * (21.0 * n00.x + p12s5 * RE2(0) + 8.0 * RE4(0))) / s5;
    // entries REAL(Z)
    u[2] = p3s6 * RE2(-1) + (-p3s6) * RE2(1) + -3.0 * RE4(-1) + 3.0 * RE4(1) + s7 * RE4(-3) + (-s7) * RE4(3);
    u[7] = s6 * RE2(-1) + (-s6) * RE2(1) + -1.0 * RE4(-1) + RE4(1) + (-s7) * RE4(-3) + s7 * RE4(3);
    u[9] = p3s6 * RE2(-1) + (-p3s6) * RE2(1) + 4.0 * RE4(-1) + -4.0 * RE4(1);
    // entries REAL((0,c) * Z) = -c * Im(Z)
    u[1] = -(s2 * (m3s3 * IM2(-2) + p3_15 * IM2(2) + IM4(-2) + -1.0 * IM4(2) + (-s7) * IM4(-4) + s7 * IM4(4)));
    u[6] = -(s2 * (m3s3 * IM2(-2) + p3_15 * IM2(2) + IM4(-2) + -1.0 * IM4(2) + s7 * IM4(-4) + (-s7) * IM4(4)));
    u[8] = -(s2 * ((-s3) * IM2(-2) + s3 * IM2(2) + -2.0 * IM4(-2) + 2.0 * IM4(2)));
    u[4] = -((-s6) * IM2(-1) + (-s6) * IM2(1) + IM4(-1) + IM4(1) + (-s7) * IM4(-3) + (-s7) * IM4(3));
    u[11] = -((-p3s6) * IM2(-1) + (-p3s6) * IM2(1) + 3.0 * IM4(-1) + 3.0 * IM4(1) + s7 * IM4(-3) + s7 * IM4(3));
    u[13] = -((-p3s6) * IM2(-1) + (-p3s6) * IM2(1) + 4.0 * (-1.0 * IM4(-1) + -1.0 * IM4(1)));
#undef RE2
#undef IM2
#undef RE4
#undef IM4
    const double k = 0x1.35370ba079db2p-5;        // Sqrt(Pi/5.)/21.
    const double c0 = 0x1.c5bf891b4ef6ap+1 * n00.x;  // f_ev_c0
#pragma unroll
    for (int q = 0; q < 15; ++q) u[q] = u[q] * k / c0;
}

// index of the sorted quadruple (a<=b<=c<=d), values 0..2, in the 15-entry list above
__host__ __device__ __forceinline__ int a4_unique_index(int a, int b, int c, int d) {
    // sort 4 small ints
    int t;
#define SW(x, y) if (x > y) { t = x; x = y; y = t; }
    SW(a, b) SW(c, d) SW(a, c) SW(b, d) SW(b, c)
#undef SW
    const int code = a * 27 + b * 9 + c * 3 + d;
    switch (code) {
        case 0: return 0;  case 1: return 1;  case 2: return 2;  case 4: return 3;  case 5: return 4;
        case 8: return 5;  case 13: return 6; case 14: return 7; case 17: return 8; case 26: return 9;
        case 40: return 10; case 41: return 11; case 44: return 12; case 53: return 13; default: return 14;
    }
}

// ---------------------------------------------------------------------------------------------
// symmetric 3x3 eigen decomposition (cyclic Jacobi), eigenvalues descending, V[:,i] eigenvectors.
// Replaces LAPACK dsyev('V','U') of src/frames.f90:75; eigenvector signs / degenerate bases are
// implementation-defined there too: here the largest |component| is made positive.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void eig3_sym(const double m[3][3], double w[3], double V[3][3]) {
    double a[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { a[i][j] = m[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[2][1] = a[1][2];   // 'U': upper triangle is the input
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double dia = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-17 * dia || off == 0.0) break;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;
            const double apq = a[p][q];
            if (apq != 0.0) {
                // tan of the rotation angle, t = sgn(theta)/(|theta| + sqrt(theta^2+1)) with theta = (aqq-app)/(2 apq), written
                // with ONE division and one square root: t = sgn(d) apq / (|d| + sqrt(d^2 + apq^2)), d = (aqq-app)/2
                const double dlt = 0.5 * (a[q][q] - a[p][p]);
                const double t = (dlt >= 0 ? apq : -apq) / (fabs(dlt) + sqrt(fma(dlt, dlt, apq * apq)));
                const double c = rsqrt(fma(t, t, 1.0)), s = t * c;
                const int k = 3 - p - q;
                const double akp = a[k][p], akq = a[k][q];
                a[p][p] -= t * apq;
                a[q][q] += t * apq;
                a[p][q] = a[q][p] = 0.0;
                a[k][p] = a[p][k] = c * akp - s * akq;
                a[k][q] = a[q][k] = s * akp + c * akq;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double vip = V[i][p], viq = V[i][q];
                    V[i][p] = c * vip - s * viq;
                    V[i][q] = s * vip + c * viq;
                }
            }
        }
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
    // sort descending (largest eigenvalue first, src/frames.f90:76-79)
#define SWAPCOL(i, j)                                                     \
    if (w[i] < w[j]) {                                                    \
        double t = w[i]; w[i] = w[j]; w[j] = t;                           \
        for (int r = 0; r < 3; ++r) { t = V[r][i]; V[r][i] = V[r][j]; V[r][j] = t; } \
    }
    SWAPCOL(0, 1) SWAPCOL(0, 2) SWAPCOL(1, 2)
#undef SWAPCOL
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int im = 0;
        if (fabs(V[1][i]) > fabs(V[im][i])) im = 1;
        if (fabs(V[2][i]) > fabs(V[im][i])) im = 2;
        if (V[im][i] < 0) { V[0][i] = -V[0][i]; V[1][i] = -V[1][i]; V[2][i] = -V[2][i]; }
    }
}

// eigframe(M, plane): src/frames.f90:24-60.  plane: 0 'ij', 1 'xy', 2 'xz'.  ei[i][:] = i-th eigenvector.
__device__ __forceinline__ void eigframe(const double m[3][3], int plane, double ei[3][3], double lam[3]) {
    double V[3][3], w[3];
    eig3_sym(m, w, V);
    int sort[3] = {0, 1, 2};
    if (plane != 0) {
        const int k = (plane == 1) ? 2 : 1;
        int imax = 0;
        if (fabs(V[k][1]) > fabs(V[k][imax])) imax = 1;
        if (fabs(V[k][2]) > fabs(V[k][imax])) imax = 2;
        if (imax == 0) { sort[0] = 1; sort[1] = 2; sort[2] = 0; }
        if (imax == 1) { sort[0] = 0; sort[1] = 2; sort[2] = 1; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        lam[i] = w[sort[i]];
#pragma unroll
        for (int x = 0; x < 3; ++x) ei[i][x] = V[x][sort[i]];
    }
}

// ---- node loads / a2 shared by the field kernels
// red = 0: full nlm array (n_2^m at 0-based row 3+m, n_4^m at 10+m); red = 1: reduced form rnlm (rows m >= 0 only,
// src/reducedform.f90:160-187: n_2^m at row 1+m, n_4^m at 4+m)
__device__ __forceinline__ void load_m_ge0(const double2* __restrict__ nlm, long long ld, long long p,
                                           double2& n00, double2 n2[3], double2 n4[5], int red = 0) {
    const int r2 = red ? 1 : 3, r4 = red ? 4 : 10;
    n00 = nlm[p];
#pragma unroll
    for (int m = 0; m < 3; ++m) n2[m] = nlm[(long long)(r2 + m) * ld + p];
#pragma unroll
    for (int m = 0; m < 5; ++m) n4[m] = nlm[(long long)(r4 + m) * ld + p];
}

__device__ __forceinline__ void a2_from(double2 n00, const double2 n2[3], double a[3][3]) {
    double a2v[6];
    double2 h[3];
    sfb::ev_c2_mandel(n00, n2[0], n2[1], n2[2], a2v, h);
    // src/moments.f90:37-44 returns f_ev_c2 directly (no Mandel round trip): undo the sqrt(2) scaling exactly
    // by forming the off-diagonals from the same quotients
    const double2 h1 = h[1], h2 = h[2];
    const double s215 = 0.3651483716701107;
    a[0][0] = a2v[0]; a[1][1] = a2v[1]; a[2][2] = a2v[2];
    a[0][1] = a[1][0] = s215 * (-h2.y);
    a[0][2] = a[2][0] = s215 * (-h1.x);
    a[1][2] = a[2][1] = s215 * (h1.y);
}

// ---------------------------------------------------------------------------------------------
// Eij_tranisotropic, n'=1
// ---------------------------------------------------------------------------------------------
struct EijCoef {          // host-evaluated scalars (libm pow like the reference): src/rheologies.f90:123-135
    double sA, sB, sC;    // Sachs  (ef = +1)
    double tA, tB, tC;    // Taylor (ef = -1)
    double s_iso;         // 1 + 2/15*sB + 2/3*sC        src/homogenizations.f90:206
    double t_iso;         // 1 + 2/15*tB + 2/3*tC        src/homogenizations.f90:235
    double alpha;
};

__device__ __forceinline__ void mat_to_vec(const double t[3][3], double v[6]) {
    v[0] = t[0][0]; v[1] = t[1][1]; v[2] = t[2][2];
    v[3] = SFB_SQRT2 * t[1][2]; v[4] = SFB_SQRT2 * t[0][2]; v[5] = SFB_SQRT2 * t[0][1];
}
__device__ __forceinline__ void vec_to_mat(const double v[6], double t[3][3]) {
    t[0][0] = v[0]; t[1][1] = v[1]; t[2][2] = v[2];
    // v/sqrt(2) as a multiplication by the correctly rounded reciprocal (<= 1 ulp from the reference's division)
    const double r2 = 0.7071067811865476;
    t[0][1] = t[1][0] = v[5] * r2; t[0][2] = t[2][0] = v[4] * r2; t[1][2] = t[2][1] = v[3] * r2;
}
__device__ __forceinline__ double dinner22(const double A[3][3], const double B[3][3]) {   // A_ij B_ji
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) r += A[i][j] * B[j][i];
        s += r;
    }
    return s;
}

// unblocked left-looking Cholesky of the LOWER triangle (LAPACK dpotf2 'L' semantics, which is what
// dposv of the reference's LAPACK leaves behind on failure: failed pivot stored, info = column)
__device__ __forceinline__ int potf2_lower(double a[6][6], double* invd = nullptr) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double ajj = a[j][j];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < j) ajj -= a[j][k] * a[j][k];
        if (!(ajj > 0.0)) { a[j][j] = ajj; return j + 1; }
        ajj = sqrt(ajj);
        a[j][j] = ajj;
        // one reciprocal per column instead of a division per entry (<= 1 ulp from LAPACK's x/ajj; the Eij tolerance is 1e-12)
        const double rj = 1.0 / ajj;
        if (invd) invd[j] = rj;
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i > j) {
            double s = a[i][j];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < j) s -= a[i][k] * a[j][k];
            a[i][j] = s * rj;
        }
    }
    return 0;
}
// triangular solves with the reciprocal pivots computed once per node (6 divisions instead of 12 per solve)
__device__ __forceinline__ void potrs_lower(const double a[6][6], const double invd[6], double x[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < i) s -= a[i][k] * x[k];
        x[i] = s * invd[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k > i) s -= a[k][i] * x[k];
        x[i] = s * invd[i];
    }
}

// Taylor matrix P of src/homogenizations.f90:165-170 (full, non-symmetric)
__device__ __forceinline__ void taylor_P(const double a2v[6], const double a4p[21], const EijCoef& K, double P[6][6]) {
    double a2m[3][3];
    vec_to_mat(a2v, a2m);
    const double s = SFB_SQRT2;
    const double Lm[6][6] = {
        {2 * a2m[0][0], 0.0, 0.0, 0.0, s * a2m[0][2], s * a2m[0][1]},
        {0.0, 2 * a2m[1][1], 0.0, s * a2m[1][2], 0.0, s * a2m[0][1]},
        {0.0, 0.0, 2 * a2m[2][2], s * a2m[1][2], s * a2m[0][2], 0.0},
        {0.0, s * a2m[1][2], s * a2m[1][2], a2m[1][1] + a2m[2][2], a2m[0][1], a2m[0][2]},
        {s * a2m[0][2], 0.0, s * a2m[0][2], a2m[0][1], a2m[0][0] + a2m[2][2], a2m[1][2]},
        {s * a2m[0][1], s * a2m[0][1], 0.0, a2m[0][2], a2m[1][2], a2m[0][0] + a2m[1][1]}};
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double idv = (i < 3) ? 1.0 : 0.0;
            P[i][j] = ((i == j ? 1.0 : 0.0) - K.tA * (idv * a2v[j])) + K.tB * a4p[i <= j ? tri6(i, j) : tri6(j, i)] + K.tC * Lm[i][j];
        }
}

// The reference's ill-posed branch (src/homogenizations.f90:177-185): P_reg = P^T P + 1e-6 I built from the PARTIALLY
// FACTORISED P that dposv leaves behind, rhs P^T tau.  Rare (unphysical states) and register hungry, so it is kept out
// of line and redoes the factorisation; returns the status flags, x <- solution for the right-hand side tv.
static __device__ __noinline__ int taylor_fallback_solve(const double a2v[6], const double a4p[21], const EijCoef& K, const double tv[6],
                                                  double x[6]) {
    double F[6][6];
    taylor_P(a2v, a4p, K, F);
    potf2_lower(F);
    double R[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s += F[k][i] * F[k][j];
            R[i][j] = s + (i == j ? 0x1.0c6f7ap-20 : 0.0);      // 1e-6 is a real(4) literal
        }
    int status = SFB_ST_TAYLOR_FALLBACK;
    if (potf2_lower(R) != 0) status |= SFB_ST_TAYLOR_FAILED;
    double invd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) invd[i] = 1.0 / R[i][i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += F[k][i] * tv[k];
        x[i] = s;
    }
    potrs_lower(R, invd, x);
    return status;
}

// Eij = (E11,E22,E33,E23,E13,E12) w.r.t. the rows of e[3][3].  returns status flags.
// SACHS_GIVEN: the six Sachs ratios come from the caller (n'=3 closure, sfb_moments_hi.cuh); the Taylor part is
// always the n'=1 solve with the caller's coefficients (src/homogenizations.f90:148).
template <bool SACHS_GIVEN = false>
__device__ __forceinline__ int eij_tranisotropic(double2 n00, const double2 n2[3], const double2 n4[5],
                                                 const double e[3][3], const EijCoef& K, double E[6],
                                                 const double* Es_given = nullptr) {
    double a2v[6], a4p[21];
    ev_c2_mandel(n00, n2[0], n2[1], n2[2], a2v);
    ev_c4_mandel(n00, n2, n4, a4p);
    double a2m[3][3];
    vec_to_mat(a2v, a2m);
    // ---- Cholesky factor of the Taylor matrix: dposv('L') reads and writes the lower triangle only, so only that
    // half is ever touched here (the unrolled code keeps 21 entries in registers, not 36)
    int status = 0;
    double F[6][6];
    {
        double P[6][6];
        taylor_P(a2v, a4p, K, P);
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (j <= i) F[i][j] = P[i][j];
    }
    double invd[6];
    const int info = potf2_lower(F, invd);
    const double inv_t_iso = 1.0 / K.t_iso;
    bool finite = true;
#ifdef SFB_EIJ_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int q = 0; q < 6; ++q) {
        // (v,w) pairs: 11,22,33,23,13,12   src/enhancementfactors.f90:36-44
        const int iv = (q < 3) ? q : (q == 3 ? 1 : 0);
        const int iw = (q < 3) ? q : (q == 5 ? 1 : 2);
        double ev[3], ew[3];       // rows iv, iw of the frame, selected without local-memory indexing
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ev[i] = iv == 0 ? e[0][i] : (iv == 1 ? e[1][i] : e[2][i]);
            ew[i] = iw == 0 ? e[0][i] : (iw == 1 ? e[1][i] : e[2][i]);
        }
        double tau[3][3], vw[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                vw[i][j] = ev[i] * ew[j];
                tau[i][j] = q < 3 ? ((i == j) ? 1.0 / 3.0 : 0.0) - ev[i] * ev[j]      // tau_vv
                                  : ev[i] * ew[j] + ew[i] * ev[j];                    // tau_vw
            }
        double tv[6];
        mat_to_vec(tau, tv);
        double Es;
        if (SACHS_GIVEN) {
            Es = Es_given[q];
        } else {
        // ---- Sachs (src/homogenizations.f90:88-91,115)
        double a4t_v[6], a4t[3][3];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) s += a4p[i <= j ? tri6(i, j) : tri6(j, i)] * tv[j];
            a4t_v[i] = s;
        }
        vec_to_mat(a4t_v, a4t);
        const double a2tau = dinner22(a2m, tau);
        double eps[3][3], epsi[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double ac = 0.0;   // tau.a2 + a2.tau
#pragma unroll
                for (int k = 0; k < 3; ++k) ac += tau[i][k] * a2m[k][j];
                double ac2 = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) ac2 += a2m[i][k] * tau[k][j];
                eps[i][j] = ((1.0 * tau[i][j] - K.sA * a2tau * (i == j ? 1.0 : 0.0)) + K.sB * a4t[i][j]) + K.sC * (ac + ac2);
                epsi[i][j] = K.s_iso * tau[i][j];
            }
        Es = dinner22(eps, vw) / dinner22(epsi, vw);
        }
        // ---- Taylor (src/homogenizations.f90:172-188)
        double x[6];
        if (info == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) x[i] = tv[i];
            potrs_lower(F, invd, x);
        } else {      // cold path: hand COPIES to the out-of-line routine so that the hot arrays stay in registers
            double a2c[6], a4c[21], tvc[6], xc[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { a2c[i] = a2v[i]; tvc[i] = tv[i]; }
#pragma unroll
            for (int i = 0; i < 21; ++i) a4c[i] = a4p[i];
            status |= taylor_fallback_solve(a2c, a4c, K, tvc, xc);
#pragma unroll
            for (int i = 0; i < 6; ++i) x[i] = xc[i];
        }
        double et[3][3], eti[3][3];
        vec_to_mat(x, et);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) eti[i][j] = tau[i][j] * inv_t_iso;
        const double Et = dinner22(et, vw) / dinner22(eti, vw);
        const double Eq = (1 - K.alpha) * Es + K.alpha * Et;
#pragma unroll
        for (int r = 0; r < 6; ++r)
            if (r == q) E[r] = Eq;
        finite = finite && isfinite(Eq);
    }
    if (!finite) status |= SFB_ST_NONFINITE;
    return status;
}

// Evw_tranisotropic (src/enhancementfactors.f90:47-69): ONE generalized enhancement factor for an arbitrary pair (v, w) and
// an arbitrary stress tau, (1 - alpha) * Sachs + alpha * Taylor, each as the ratio to the isotropic response.  Same algebra
// as one pass of the loop in eij_tranisotropic (kept separate so that the tuned six-pair kernel is not perturbed).
__device__ __forceinline__ int evw_tranisotropic(double2 n00, const double2 n2[3], const double2 n4[5], const double v[3], const double w[3],
                                                 const double tau[3][3], const EijCoef& K, double& Evw) {
    double a2v[6], a4p[21], a2m[3][3];
    ev_c2_mandel(n00, n2[0], n2[1], n2[2], a2v);
    ev_c4_mandel(n00, n2, n4, a4p);
    vec_to_mat(a2v, a2m);
    double vw[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vw[i][j] = v[i] * w[j];
    double tv[6];
    mat_to_vec(tau, tv);
    // ---- Sachs (src/homogenizations.f90:88-91,115)
    double a4t_v[6], a4t[3][3];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) s += a4p[i <= j ? tri6(i, j) : tri6(j, i)] * tv[j];
        a4t_v[i] = s;
    }
    vec_to_mat(a4t_v, a4t);
    const double a2tau = dinner22(a2m, tau);
    double eps[3][3], epsi[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double ac = 0.0, ac2 = 0.0;   // tau.a2 + a2.tau
#pragma unroll
            for (int k = 0; k < 3; ++k) { ac += tau[i][k] * a2m[k][j]; ac2 += a2m[i][k] * tau[k][j]; }
            eps[i][j] = ((1.0 * tau[i][j] - K.sA * a2tau * (i == j ? 1.0 : 0.0)) + K.sB * a4t[i][j]) + K.sC * (ac + ac2);
            epsi[i][j] = K.s_iso * tau[i][j];
        }
    const double Es = dinner22(eps, vw) / dinner22(epsi, vw);
    // ---- Taylor (src/homogenizations.f90:172-188)
    int status = 0;
    double F[6][6], invd[6];
    {
        double P[6][6];
        taylor_P(a2v, a4p, K, P);
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (j <= i) F[i][j] = P[i][j];
    }
    double x[6];
    if (potf2_lower(F, invd) == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = tv[i];
        potrs_lower(F, invd, x);
    } else {
        status |= taylor_fallback_solve(a2v, a4p, K, tv, x);
    }
    double et[3][3], eti[3][3];
    vec_to_mat(x, et);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) eti[i][j] = tau[i][j] / K.t_iso;
    const double Et = dinner22(et, vw) / dinner22(eti, vw);
    Evw = (1 - K.alpha) * Es + K.alpha * Et;
    if (!isfinite(Evw)) status |= SFB_ST_NONFINITE;
    return status;
}

}  // namespace sfb
