/* CPU oracle, C restatement of the reference's DENSE per-node algorithm.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (never linked into libspecfab_b200.so).  It keeps the loop
 * structure of the reference -- build the full n x n operator from the Gaunt tables for every
 * node and every step, then a dense matvec (src/dynamics.f90:94-96, 295-297, 108) -- so it doubles
 * as the timed CPU baseline ("kind": "port"; gfortran is not available in this image).
 * Validated against oracle/specfab_oracle.py (tests/test_oracle_c.py).  Parity status: see the
 * header of specfab_oracle.py ("parity unpinned" w.r.t. a compiled reference).
 *
 * Build: gcc -O2 -std=c11 -fcx-fortran-rules -fopenmp -shared -fPIC -o oracle/_build/liboracle.so oracle/specfab_oracle.c -lm
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 231
#define NCAT 15
typedef double complex cplx;

static int g_L = 0, g_n = 0;
/* tables stored [i][j][k] (k fastest); the reference's are Fortran (i,j,k) column-major */
static double *GC, *GCm, *GC_m1, *GC_p1;
static double g_nu, g_expo;
static double g_regdiag[NMAX], g_ldiag[NMAX];
static const double Pi = 3.141592653589793; /* src/header.f90:12 */

static const double REG_EXPO[9] = {1.700, 1.150, 1.600, 2.000, 2.000, 2.000, 2.500, 2.500, 3.000};
static const double REG_NU[9] = {1.9879322126397958, 3.0011508426238862, 5.7498069921352384, 10.7048905312159288,
                                 10.6068117205577668, 13.3591023418822363, 15.3094482670021108, 16.4844589176829217,
                                 19.9467342880730136}; /* src/include/regcalib.f90:1-36 */

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* tables: dense [231][231][15] doubles, C order (i,j,k) */
int orc_init(int L, const double* gc, const double* gcm, const double* gcm1, const double* gcp1) {
    if (L % 2 || L < 4 || L > 20) return -1;
    g_L = L;
    g_n = (L + 1) * (L + 2) / 2;
    size_t sz = (size_t)NMAX * NMAX * NCAT * sizeof(double);
    if (!GC) { GC = malloc(sz); GCm = malloc(sz); GC_m1 = malloc(sz); GC_p1 = malloc(sz); }
    memcpy(GC, gc, sz); memcpy(GCm, gcm, sz); memcpy(GC_m1, gcm1, sz); memcpy(GC_p1, gcp1, sz);
    g_expo = REG_EXPO[L / 2 - 2];
    g_nu = REG_NU[L / 2 - 2];
    int j = 0;
    for (int l = 0; l <= L; l += 2)
        for (int m = -l; m <= l; ++m, ++j) {
            g_ldiag[j] = -(double)(l * (l + 1));                                   /* src/dynamics.f90:22 */
            g_regdiag[j] = pow(fabs(g_ldiag[j] / (double)(L * (L + 1))), g_expo); /* src/dynamics.f90:512 */
        }
    return g_n;
}

/* src/dynamics.f90:563-579 */
static void quad_rr(const double M[3][3], cplx q[5]) {
    const double fsq = sqrt(2 * Pi / 15);
    const float twothirds = 2.f / 3; /* 2./3 is real(4) */
    q[0] = fsq * (M[0][0] - M[1][1] + 2 * I * M[0][1]);
    q[1] = 2 * fsq * (M[0][2] + I * M[1][2]);
    q[2] = -((double)twothirds * sqrt(Pi / 5)) * (M[0][0] + M[1][1] - 2 * M[2][2]);
    q[3] = -2 * fsq * (M[0][2] - I * M[1][2]);
    q[4] = fsq * (M[0][0] - M[1][1] - 2 * I * M[0][1]);
}

/* src/dynamics.f90:581-593 */
static void quad_tp(const double M[3][3], cplx q[3]) {
    const double fsq1 = sqrt(2 * Pi / 3);
    const float s2 = sqrtf(2.f);
    q[0] = fsq1 * (M[1][2] - I * M[0][2]);
    q[1] = fsq1 * ((double)s2 * M[0][1]);
    q[2] = fsq1 * (-M[1][2] - I * M[0][2]);
}

/* src/dynamics.f90:52-97: M += M_LROT(eps, omg, iota, zeta)   (M is n x n, row-major) */
static void add_M_LROT(cplx* M, const double eps[3][3], const double omg[3][3], double iota, double zeta) {
    const int n = g_n;
    double sq[3][3], E[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) sq[i][j] = eps[i][0] * eps[0][j] + eps[i][1] * eps[1][j] + eps[i][2] * eps[2][j];
    const double zetanorm = zeta / sqrt(sq[0][0] + sq[1][1] + sq[2][2]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) E[i][j] = iota * eps[i][j] + zetanorm * sq[i][j];
    cplx qe[5], qo[3];
    quad_rr(E, qe);
    quad_tp(omg, qo);
    const double s3 = (double)sqrtf(3.f), s6 = (double)sqrtf(6.f);
    const double s56 = (double)sqrtf(5.f / 6), s23 = (double)sqrtf(2.f / 3), s32 = (double)sqrtf(3.f / 2);
    cplx g0[6], gz[6], gn[6], gp[6];
    g0[0] = 0; for (int k = 0; k < 5; ++k) g0[k + 1] = 3 * qe[k];
    gz[0] = -(I * s3) * qo[1]; gz[1] = -qe[0]; gz[2] = gz[3] = gz[4] = 0; gz[5] = qe[4];
    gn[0] = -((I * 6) / s6) * qo[0] + s56 * qe[1]; gn[1] = 0; gn[2] = qe[0]; gn[3] = s23 * qe[1]; gn[4] = s32 * qe[2]; gn[5] = 2 * qe[3];
    gp[0] = +((I * 6) / s6) * qo[2] + s56 * qe[3]; gp[1] = 2 * qe[1]; gp[2] = s32 * qe[2]; gp[3] = s23 * qe[3]; gp[4] = qe[4]; gp[5] = 0;
    for (int ii = 0; ii < n; ++ii)
        for (int jj = 0; jj < n; ++jj) {
            const size_t o = ((size_t)ii * NMAX + jj) * NCAT;
            cplx a = 0, b = 0, c = 0, d = 0;
            for (int k = 0; k < 6; ++k) {
                a += GC[o + k] * g0[k]; b += GCm[o + k] * gz[k]; c += GC_m1[o + k] * gn[k]; d += GC_p1[o + k] * gp[k];
            }
            M[(size_t)ii * n + jj] += -1 * (a + b + c + d);
        }
}

/* src/include/ddrx-coupling-weights.f90 + src/dynamics.f90:291-293 */
static void ddrx_weights(const double tau[3][3], cplx g[15]) {
    cplx qt[5];
    quad_rr(tau, qt);
    const cplx qm2 = qt[0], qm1 = qt[1], q0 = qt[2], qp1 = qt[3], qp2 = qt[4];
    const double s5 = (double)sqrtf(5.f), s15 = (double)sqrtf(1.5f), s6 = (double)sqrtf(6.f), s2 = (double)sqrtf(2.f), s3 = (double)sqrtf(3.f);
    const double ms6 = (double)(-1.0f * sqrtf(6.f)), c2s14 = (double)(2 * sqrtf(14.f)), c4s7 = (double)(4 * sqrtf(7.f)), c3s5 = (double)(3.f * sqrtf(5.f));
    const double k = (3 * sqrt(5 / Pi)) / 28.0;
    g[0] = (7.0 * (q0 * q0 + (-2.0 * qm1) * qp1 + (2.0 * qm2) * qp2)) / s5;
    g[1] = s15 * (qm1 * qm1) + (-2.0 * q0) * qm2;
    g[2] = q0 * qm1 + (ms6 * qp1) * qm2;
    g[3] = q0 * q0 + (-1.0 * qm1) * qp1 + (-2.0 * qm2) * qp2;
    g[4] = q0 * qp1 + (ms6 * qm1) * qp2;
    g[5] = s15 * (qp1 * qp1) + (-2.0 * q0) * qp2;
    g[6] = (-(c2s14 * (qm2 * qm2))) / 3.0;
    g[7] = (-((c4s7 * qm1) * qm2)) / 3.0;
    g[8] = (-(4 * (s2 * (qm1 * qm1) + (s3 * q0) * qm2))) / 3.0;
    g[9] = (-(4 * ((s6 * q0) * qm1 + qp1 * qm2))) / 3.0;
    g[10] = (-(4 * (3.0 * (q0 * q0) + (4.0 * qm1) * qp1 + qm2 * qp2))) / c3s5;
    g[11] = (-(4 * ((s6 * q0) * qp1 + qm1 * qp2))) / 3.0;
    g[12] = (-(4 * (s2 * (qp1 * qp1) + (s3 * q0) * qp2))) / 3.0;
    g[13] = (-((c4s7 * qp1) * qp2)) / 3.0;
    g[14] = (-(c2s14 * (qp2 * qp2))) / 3.0;
    double dd = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) dd += tau[i][j] * tau[j][i];
    for (int q = 0; q < 15; ++q) g[q] = ((k * g[q]) * 5) / dd;
}

/* a2 / a4 in Mandel form: src/moments.f90:164-218, ev_c2__body.f90, ev_c4_Mandel__body.f90 */
static void ev_ck_mandel(const cplx* nlm, double a2v[6], double e[6][6]) {
    const cplx n00 = nlm[0];
    const cplx h0 = nlm[3] / n00, h1 = nlm[4] / n00, h2 = nlm[5] / n00;
    const double c = 0.5 * sqrt(2.0 / 3), s215 = sqrt(2 / 15.0), s = sqrt(2.0);
    double ev[3][3];
    ev[0][0] = -(c * creal(h0)) + creal(h2);
    ev[1][1] = -(c * creal(h0)) - creal(h2);
    ev[2][2] = sqrt(2.0 / 3) * creal(h0);
    ev[0][1] = ev[1][0] = -cimag(h2);
    ev[0][2] = ev[2][0] = -creal(h1);
    ev[1][2] = ev[2][1] = cimag(h1);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ev[i][j] = s215 * ev[i][j] + (i == j ? 1.0 / 3.0 : 0.0);
    a2v[0] = ev[0][0]; a2v[1] = ev[1][1]; a2v[2] = ev[2][2]; a2v[3] = s * ev[1][2]; a2v[4] = s * ev[0][2]; a2v[5] = s * ev[0][1];
    const double s5 = sqrt(5.0), s6 = sqrt(6.0), s7 = sqrt(7.0), s10 = sqrt(10.0), s30 = sqrt(30.0), s70 = sqrt(70.0), s15 = sqrt(15.0), s3 = sqrt(3.0);
    const double r00 = creal(n00), r20 = creal(nlm[3]), r21 = creal(nlm[4]), r22 = creal(nlm[5]), i21 = cimag(nlm[4]), i22 = cimag(nlm[5]);
    const double r40 = creal(nlm[10]), r41 = creal(nlm[11]), r42 = creal(nlm[12]), r43 = creal(nlm[13]), r44 = creal(nlm[14]);
    const double i41 = cimag(nlm[11]), i42 = cimag(nlm[12]), i43 = cimag(nlm[13]), i44 = cimag(nlm[14]);
    e[0][0] = 21.0 * r00 + (s5 * -6.0) * r20 + (s30 * 6.0) * r22 + 3.0 * r40 + (s10 * -2.0) * r42 + s70 * r44;
    e[0][1] = 7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44;
    e[0][2] = 7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42;
    e[0][3] = s10 * (s6 * i21 + -1.0 * i41 + s7 * i43);
    e[0][4] = (s10 * -1.0) * ((3.0 * s6) * r21 + -3.0 * r41 + s7 * r43);
    e[0][5] = (-2.0 * s5) * (pow(3.0, 1.5) * i22 + -1.0 * i42 + s7 * i44);
    e[1][1] = 21.0 * r00 + (s5 * -6.0) * (r20 + s6 * r22) + 3.0 * r40 + s10 * (2.0 * r42 + s7 * r44);
    e[1][2] = 7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42);
    e[1][3] = s10 * ((3.0 * s6) * i21 + -3.0 * i41 + (-1.0 * s7) * i43);
    e[1][4] = (s15 * -2.0) * r21 + s10 * (r41 + s7 * r43);
    e[1][5] = (2.0 * s5) * ((-3.0 * s3) * i22 + i42 + s7 * i44);
    e[2][2] = 21.0 * r00 + (12.0 * s5) * r20 + 8.0 * r40;
    e[2][3] = s10 * ((3.0 * s6) * i21 + 4.0 * i41);
    e[2][4] = (s10 * -1.0) * ((3.0 * s6) * r21 + 4.0 * r41);
    e[2][5] = (-2.0 * s5) * (s3 * i22 + 2.0 * i42);
    e[3][3] = 2.0 * (7.0 * r00 + s5 * (r20 + (-1.0 * s6) * r22) + -2.0 * (2.0 * r40 + s10 * r42));
    e[3][4] = (s10 * -2.0) * (s3 * i22 + 2.0 * i42);
    e[3][5] = s5 * ((-2.0 * s6) * r21 + 2.0 * (r41 + s7 * r43));
    e[4][4] = 2.0 * (7.0 * r00 + s5 * (r20 + s6 * r22) + -4.0 * r40 + (s10 * 2.0) * r42);
    e[4][5] = (2.0 * s5) * (s6 * i21 + -1.0 * i41 + s7 * i43);
    e[5][5] = 2.0 * (7.0 * r00 + (-2.0 * s5) * r20 + r40 + (-1.0 * s70) * r44);
    const double k = (2 * sqrt(Pi)) / 105.0, c0 = sqrt(4 * Pi) * r00;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { e[a][b] = e[a][b] * k / c0; e[b][a] = e[a][b]; }
}

/* src/dynamics.f90:402-422 */
static double ev_D2(const cplx* nlm, const double tau[3][3]) {
    const double s = sqrt(2.0);
    double sq[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) sq[i][j] = tau[i][0] * tau[0][j] + tau[i][1] * tau[1][j] + tau[i][2] * tau[2][j];
    const double tv[6] = {tau[0][0], tau[1][1], tau[2][2], s * tau[1][2], s * tau[0][2], s * tau[0][1]};
    const double sv[6] = {sq[0][0], sq[1][1], sq[2][2], s * sq[1][2], s * sq[0][2], s * sq[0][1]};
    const double norm = sq[0][0] + sq[1][1] + sq[2][2];
    double a2v[6], e[6][6];
    ev_ck_mandel(nlm, a2v, e);
    double d1 = 0, d2 = 0;
    for (int p = 0; p < 6; ++p) d1 += sv[p] * a2v[p];
    for (int p = 0; p < 6; ++p) {
        double r = 0;
        for (int q = 0; q < 6; ++q) r += e[p][q] * tv[q];
        d2 += tv[p] * r;
    }
    return 5 * (d1 - d2) / norm;
}

typedef struct {
    double dt, iota, zeta, nu_mult, gamma0, lambda;
    int use_lrot, use_ddrx, use_cdrx, use_reg, rk4;
} orc_opts;

/* dn/dt = M n with M assembled densely per call, as the reference's callers do */
static void rhs(const cplx* y, const double ug[3][3], const double tau[3][3], const orc_opts* o, cplx* M, cplx* k) {
    const int n = g_n;
    double D[3][3], W[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { D[i][j] = (ug[i][j] + ug[j][i]) / 2; W[i][j] = (ug[i][j] - ug[j][i]) / 2; }
    memset(M, 0, sizeof(cplx) * n * n);
    if (o->use_lrot) add_M_LROT(M, D, W, o->iota, o->zeta);
    if (o->use_ddrx) {
        cplx g[15];
        ddrx_weights(tau, g);
        const double davg = ev_D2(y, tau);
        for (int ii = 0; ii < n; ++ii) {
            for (int jj = 0; jj < n; ++jj) {
                const size_t ofs = ((size_t)ii * NMAX + jj) * NCAT;
                cplx a = 0;
                for (int q = 0; q < 15; ++q) a += GC[ofs + q] * g[q];          /* src/dynamics.f90:295-297 */
                if (ii == jj) a -= davg;                                      /* src/dynamics.f90:270-274 */
                M[(size_t)ii * n + jj] += o->gamma0 * a;
            }
        }
    }
    if (o->use_cdrx)
        for (int ii = 0; ii < n; ++ii) M[(size_t)ii * n + ii] += o->lambda * g_ldiag[ii];
    if (o->use_reg) {
        double fro = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) fro += D[i][j] * D[i][j];
        const double ratemag = g_nu * sqrt(fro);
        for (int ii = 0; ii < n; ++ii) M[(size_t)ii * n + ii] += o->nu_mult * (-ratemag * g_regdiag[ii]);
    }
    for (int ii = 0; ii < n; ++ii) {
        cplx a = 0;
        for (int jj = 0; jj < n; ++jj) a += M[(size_t)ii * n + jj] * y[jj];
        k[ii] = a;
    }
}

/* nlm: [N][n] complex (node-major, C order), ugrad/tau: [N][3][3]; nsteps steps per node; returns 0 */
int orc_step_batch(double* nlm_ri, int64_t N, const double* ugrad, const double* tau, const orc_opts* o, int nsteps) {
    const int n = g_n;
    if (!n) return -1;
#pragma omp parallel
    {
        cplx* M = malloc(sizeof(cplx) * n * n);
        cplx* k1 = malloc(sizeof(cplx) * n * 6);
        cplx *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n, *yt = k4 + n, *y = yt + n;
#pragma omp for schedule(static)
        for (int64_t p = 0; p < N; ++p) {
            cplx* x = (cplx*)(nlm_ri + 2 * p * n);
            double ug[3][3], ta[3][3];
            memcpy(ug, ugrad + 9 * p, sizeof ug);
            if (tau) memcpy(ta, tau + 9 * p, sizeof ta);
            else
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) ta[i][j] = (ug[i][j] + ug[j][i]) / 2;
            memcpy(y, x, sizeof(cplx) * n);
            for (int s = 0; s < nsteps; ++s) {
                if (!o->rk4) {
                    rhs(y, ug, ta, o, M, k1);
                    for (int i = 0; i < n; ++i) y[i] = y[i] + o->dt * k1[i];     /* src/dynamics.f90:108 */
                } else {
                    rhs(y, ug, ta, o, M, k1);
                    for (int i = 0; i < n; ++i) yt[i] = y[i] + (o->dt / 2) * k1[i];
                    rhs(yt, ug, ta, o, M, k2);
                    for (int i = 0; i < n; ++i) yt[i] = y[i] + (o->dt / 2) * k2[i];
                    rhs(yt, ug, ta, o, M, k3);
                    for (int i = 0; i < n; ++i) yt[i] = y[i] + o->dt * k3[i];
                    rhs(yt, ug, ta, o, M, k4);
                    for (int i = 0; i < n; ++i) y[i] = y[i] + (o->dt / 6) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
                }
            }
            memcpy(x, y, sizeof(cplx) * n);
        }
        free(M); free(k1);
    }
    return 0;
}
