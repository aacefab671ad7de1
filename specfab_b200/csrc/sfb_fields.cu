// Kernels + launchers for a2 / a4 / eigenframe / Eij on batches of nodes (thread per node).
#include <cstdlib>
#include "sfb_fields.cuh"

namespace {

constexpr int kBlock = 128;

using sfb::load_m_ge0;
using sfb::a2_from;

// red: the state array is in reduced form (see load_m_ge0)
__global__ void __launch_bounds__(kBlock) a2_kernel(const double2* __restrict__ nlm, long long N, long long ld,
                                                    double* __restrict__ out, long long ldo, int red) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 n00 = nlm[p], n2[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) n2[m] = nlm[(long long)((red ? 1 : 3) + m) * ld + p];
    double a[3][3];
    a2_from(n00, n2, a);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) out[(long long)(i + 3 * k) * ldo + p] = a[i][k];
}

__global__ void __launch_bounds__(kBlock) a4_kernel(const double2* __restrict__ nlm, long long N, long long ld,
                                                    double* __restrict__ out, long long ldo) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 n00 = nlm[p], n2[5], n4[9];
#pragma unroll
    for (int m = 0; m < 5; ++m) n2[m] = nlm[(long long)(1 + m) * ld + p];
#pragma unroll
    for (int m = 0; m < 9; ++m) n4[m] = nlm[(long long)(6 + m) * ld + p];
    double u[15];
    sfb::ev_c4_unique(n00, n2, n4, u);
    // a4(N,3,3,3,3) Fortran order: plane index a + 3b + 9c + 27d
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    int q = sfb::a4_unique_index(a, b, c, d);
                    // reference quirk: src/include/ev_c4__body.f90:78  ev(3,2,1,2)=ev(1,2,3,3)
                    if (a == 2 && b == 1 && c == 0 && d == 1) q = 8;
                    out[(long long)(a + 3 * b + 9 * c + 27 * d) * ldo + p] = u[q];
                }
}

// eig(nlm) (mode 0, src/frames.f90:14-22) or eigframe(M, plane) (mode 1, src/frames.f90:24-60)
__global__ void __launch_bounds__(kBlock) eig_kernel(const double2* __restrict__ nlm, const double* __restrict__ M,
                                                     long long N, long long ld, int plane,
                                                     double* __restrict__ ei, double* __restrict__ lami, long long ldo, int red) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double a[3][3];
    if (nlm) {
        double2 n00 = nlm[p], n2[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) n2[m] = nlm[(long long)((red ? 1 : 3) + m) * ld + p];
        a2_from(n00, n2, a);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) a[i][k] = M[(long long)(i + 3 * k) * ld + p];
    }
    double e[3][3], lam[3];
    sfb::eigframe(a, plane, e, lam);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        lami[(long long)i * ldo + p] = lam[i];
#pragma unroll
        for (int x = 0; x < 3; ++x) ei[(long long)(i + 3 * x) * ldo + p] = e[i][x];   // ei(N,3,3): (p,i,x)
    }
}

// Eij_tranisotropic_arr (src/specfabpy.f90:474-486).  frame given (e1,e2,e3 each (N,3) Fortran order)
// or, when e1 == nullptr, computed as the a2 eigenframe of the node (fused a2 -> eig -> Eij).
template <int MB>
__global__ void __launch_bounds__(kBlock, MB) eij_kernel(const double2* __restrict__ nlm, long long N, long long ld,
                                                     const double* __restrict__ e1, const double* __restrict__ e2,
                                                     const double* __restrict__ e3, long long lde, sfb::EijCoef K,
                                                     double* __restrict__ Eij, long long ldo,
                                                     double* __restrict__ ei_out, double* __restrict__ lam_out,
                                                     int* __restrict__ status, int red) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 n00, n2[3], n4[5];
    load_m_ge0(nlm, ld, p, n00, n2, n4, red);
    double e[3][3];
    if (e1) {
#pragma unroll
        for (int x = 0; x < 3; ++x) { e[0][x] = e1[(long long)x * lde + p]; e[1][x] = e2[(long long)x * lde + p]; e[2][x] = e3[(long long)x * lde + p]; }
    } else {
        double a[3][3], lam[3];
        a2_from(n00, n2, a);
        sfb::eigframe(a, 0, e, lam);
        if (ei_out) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                lam_out[(long long)i * ldo + p] = lam[i];
#pragma unroll
                for (int x = 0; x < 3; ++x) ei_out[(long long)(i + 3 * x) * ldo + p] = e[i][x];
            }
        }
    }
    double E[6];
    const int st = sfb::eij_tranisotropic(n00, n2, n4, e, K, E);
#pragma unroll
    for (int q = 0; q < 6; ++q) Eij[(long long)q * ldo + p] = E[q];
    if (status) status[p] = st;
}

// Evw_tranisotropic batched over nodes (src/specfabpy.f90:379-388): v, w (N,3), tau (N,3,3), one factor per node
__global__ void __launch_bounds__(kBlock) evw_kernel(const double2* __restrict__ nlm, long long N, long long ld, const double* __restrict__ v,
                                                     const double* __restrict__ w, const double* __restrict__ tau, long long lde, sfb::EijCoef K,
                                                     double* __restrict__ Evw, int* __restrict__ status) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 n00, n2[3], n4[5];
    load_m_ge0(nlm, ld, p, n00, n2, n4, 0);
    double vv[3], ww[3], t[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        vv[i] = v[(long long)i * lde + p];
        ww[i] = w[(long long)i * lde + p];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[i][k] = tau[(long long)(i + 3 * k) * lde + p];
    }
    double E;
    const int st = sfb::evw_tranisotropic(n00, n2, n4, vv, ww, t, K, E);
    Evw[p] = E;
    if (status) status[p] = st;
}

// apply_bounds (src/dynamics.f90:530-557): rescale the l=2 / l=4 blocks if their power spectrum S(l)
// (src/idealstate.f90:80-93) exceeds that of the delta function; other coefficients pass through.
__global__ void __launch_bounds__(kBlock) bounds_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long N,
                                                        long long ldi, long long ldo, int n) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 v[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) v[j] = in[(long long)j * ldi + p];
    const double S0 = v[0].x * v[0].x;                      // real(nlm(1))**2
    double s2 = 0.0, s4 = 0.0;
#pragma unroll
    for (int j = 1; j < 6; ++j) s2 += v[j].x * v[j].x + v[j].y * v[j].y;
#pragma unroll
    for (int j = 6; j < 15; ++j) s4 += v[j].x * v[j].x + v[j].y * v[j].y;
    const double S2_rel = (1.0 / 5 * s2) / S0, S4_rel = (1.0 / 9 * s4) / S0;
    const double f2 = S2_rel > 1.0 ? sqrt(S2_rel) : 1.0, f4 = S4_rel > 1.0 ? sqrt(S4_rel) : 1.0;
    out[p] = v[0];
#pragma unroll
    for (int j = 1; j < 6; ++j) out[(long long)j * ldo + p] = S2_rel > 1.0 ? make_double2(v[j].x / f2, v[j].y / f2) : v[j];
#pragma unroll
    for (int j = 6; j < 15; ++j) out[(long long)j * ldo + p] = S4_rel > 1.0 ? make_double2(v[j].x / f4, v[j].y / f4) : v[j];
    if (in != out)
        for (int j = 15; j < n; ++j) out[(long long)j * ldo + p] = in[(long long)j * ldi + p];
}

// apply_bounds on reduced-form states (what src/specfabpy/fenics/CPO.py:339-365 does per node through rnlm_to_nlm /
// apply_bounds / nlm_to_rnlm): rows 0 | 1..3 (l = 2, m = 0..2) | 4..8 (l = 4, m = 0..4).  The power sums run over m = -l..l in
// the full-form order with |n_l^-m|^2 = |n_l^m|^2, so the factors equal the full-form kernel's bit for bit.
__global__ void __launch_bounds__(kBlock) bounds_rnlm_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long N,
                                                             long long ldi, long long ldo, int r) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    double2 v[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] = in[(long long)j * ldi + p];
    const double S0 = v[0].x * v[0].x;
    double s2 = 0.0, s4 = 0.0;       // same expression shape as bounds_kernel (the compiler contracts both the same way)
#pragma unroll
    for (int m = -2; m <= 2; ++m) { const double2 w = v[1 + (m < 0 ? -m : m)]; s2 += w.x * w.x + w.y * w.y; }
#pragma unroll
    for (int m = -4; m <= 4; ++m) { const double2 w = v[4 + (m < 0 ? -m : m)]; s4 += w.x * w.x + w.y * w.y; }
    const double S2_rel = (1.0 / 5 * s2) / S0, S4_rel = (1.0 / 9 * s4) / S0;
    const double f2 = S2_rel > 1.0 ? sqrt(S2_rel) : 1.0, f4 = S4_rel > 1.0 ? sqrt(S4_rel) : 1.0;
    out[p] = v[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) out[(long long)j * ldo + p] = S2_rel > 1.0 ? make_double2(v[j].x / f2, v[j].y / f2) : v[j];
#pragma unroll
    for (int j = 4; j < 9; ++j) out[(long long)j * ldo + p] = S4_rel > 1.0 ? make_double2(v[j].x / f4, v[j].y / f4) : v[j];
    if (in != out)
        for (int j = 9; j < r; ++j) out[(long long)j * ldo + p] = in[(long long)j * ldi + p];
}

// nlm <-> rnlm (src/reducedform.f90:160-187).  rnlm rows are the m >= 0 coefficients in (l, m=0..l) order;
// rnlm_to_nlm fills n_l^{-m} = (-1)^m conj(n_l^m).  One thread per node walks the REDUCED rows of one degree group
// (blockIdx.y = l/2): each source element is read once and written to its one (to_reduced) or two (mirror) destinations,
// with all loads of the group in flight together.
__global__ void __launch_bounds__(kBlock) reduced_kernel(int to_reduced, const double2* __restrict__ src, double2* __restrict__ dst,
                                                         long long N, long long lds, long long ldd, int L) {
    const long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (p >= N) return;
    const int h = blockIdx.y, l = 2 * h;
    const int rb = h * h, fb = l * (l + 1) / 2;              // reduced row of (l, 0), full-form row of (l, 0)
    (void)L;
#pragma unroll 4
    for (int m = 0; m <= l; ++m) {
        if (to_reduced) {
            __stcs(dst + (long long)(rb + m) * ldd + p, __ldcs(src + (long long)(fb + m) * lds + p));
        } else {
            const double2 v = __ldcs(src + (long long)(rb + m) * lds + p);
            __stcs(dst + (long long)(fb + m) * ldd + p, v);
            if (m) __stcs(dst + (long long)(fb - m) * ldd + p, (m & 1) ? make_double2(-v.x, v.y) : make_double2(v.x, -v.y));
        }
    }
}

inline unsigned nblk(long long N) { return (unsigned)((N + kBlock - 1) / kBlock); }

}  // namespace

cudaError_t sfb_launch_a2(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st, int red) {
    if (N > 0) a2_kernel<<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, out, ldo, red);
    return cudaGetLastError();
}
cudaError_t sfb_launch_a4(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st) {
    if (N > 0) a4_kernel<<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, out, ldo);
    return cudaGetLastError();
}
cudaError_t sfb_launch_eig(const double2* nlm, const double* M, long long N, long long ld, int plane, double* ei, double* lami,
                           long long ldo, cudaStream_t st, int red) {
    if (N > 0) eig_kernel<<<nblk(N), kBlock, 0, st>>>(nlm, M, N, ld, plane, ei, lami, ldo, red);
    return cudaGetLastError();
}
cudaError_t sfb_launch_eij(const double2* nlm, long long N, long long ld, const double* e1, const double* e2, const double* e3,
                           long long lde, const sfb::EijCoef& K, double* Eij, long long ldo, double* ei_out, double* lam_out,
                           int* status, cudaStream_t st, int red) {
    static int mb = 0;
    // 4 CTAs of 128 threads per SM (128 registers, a few spills) beats 2 x 255 registers: the kernel is FP64-latency bound
    if (!mb) { const char* ev = getenv("SFB_EIJ_MB"); mb = ev ? atoi(ev) : 4; }
    if (N > 0) {
        if (mb == 3) eij_kernel<3><<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, e1, e2, e3, lde, K, Eij, ldo, ei_out, lam_out, status, red);
        else if (mb == 4) eij_kernel<4><<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, e1, e2, e3, lde, K, Eij, ldo, ei_out, lam_out, status, red);
        else eij_kernel<2><<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, e1, e2, e3, lde, K, Eij, ldo, ei_out, lam_out, status, red);
    }
    return cudaGetLastError();
}

cudaError_t sfb_launch_evw(const double2* nlm, long long N, long long ld, const double* v, const double* w, const double* tau, long long lde,
                           const sfb::EijCoef& K, double* Evw, int* status, cudaStream_t st) {
    if (N > 0) evw_kernel<<<nblk(N), kBlock, 0, st>>>(nlm, N, ld, v, w, tau, lde, K, Evw, status);
    return cudaGetLastError();
}

cudaError_t sfb_launch_bounds(const double2* in, double2* out, long long N, long long ldi, long long ldo, int n, cudaStream_t st) {
    if (N > 0) bounds_kernel<<<nblk(N), kBlock, 0, st>>>(in, out, N, ldi, ldo, n);
    return cudaGetLastError();
}
cudaError_t sfb_launch_bounds_rnlm(const double2* in, double2* out, long long N, long long ldi, long long ldo, int r, cudaStream_t st) {
    if (N > 0) bounds_rnlm_kernel<<<nblk(N), kBlock, 0, st>>>(in, out, N, ldi, ldo, r);
    return cudaGetLastError();
}
cudaError_t sfb_launch_reduced(int to_reduced, const double2* src, double2* dst, long long N, long long lds, long long ldd, int L,
                               cudaStream_t st) {
    if (N > 0) reduced_kernel<<<dim3(nblk(N), L / 2 + 1), kBlock, 0, st>>>(to_reduced, src, dst, N, lds, ldd, L);
    return cudaGetLastError();
}
