// Field kernels that need the 6th / 8th order structure tensors (harmonics up to l = 8), thread per node:
//   a6_arr                       src/moments.f90:57-66 (specfabpy a6, src/specfabpy.f90:601-608)
//   Eij_tranisotropic, n' = 3    src/homogenizations.f90:93-102 (Sachs) + the n'=1 Taylor solve (:148)
//   E_CAFFE_arr                  src/specfabpy.f90:543-554, src/enhancementfactors.f90:301-331, ev_D2 / ev_D4
//   pfJ                          src/idealstate.f90:111-124
// The 28 + 45 unique entries of a6 / a8 of a node live in shared memory ([entry][thread], conflict free); the
// contraction chains a8:tau:tau:tau etc. are fully unrolled over compile-time unique-entry indices.
#include <cmath>

#include "sfb_moments_hi.cuh"
#include "gen/ingest.inc"

namespace {

constexpr int kB = 64;
constexpr double kN00Iso = 0.28209479177387814;      // 1/Sqrt(4*Pi)   src/homogenizations.f90:60

// a^(k) of the isotropic state through the same code path (float32-constant round-off included), once per device
struct IsoTensors { double a4[15]; double a6[28]; double a8[45]; };
__device__ IsoTensors g_iso;

__global__ void iso_kernel() {
    const double2 z = make_double2(0.0, 0.0), n00 = make_double2(kN00Iso, 0.0);
    double2 n2[5], n4[9];
    for (int m = 0; m < 5; ++m) n2[m] = z;
    for (int m = 0; m < 9; ++m) n4[m] = z;
    sfb::ev_c4_unique(n00, n2, n4, g_iso.a4);
    auto load = [&](int j) { return j == 0 ? n00 : z; };
    ev_c6_unique(load, g_iso.a6);
    ev_c8_unique(load, g_iso.a8);
}

struct Smem {
    double a6[28][kB];
    double a8[45][kB];
};

// unique a6 (and a8) of node p into this thread's shared-memory column
template <bool WITH8>
__device__ __forceinline__ void stage_hi(Smem& s, const double2* __restrict__ nlm, long long ld, long long p) {
    auto load = [&](int j) { return nlm[(long long)j * ld + p]; };
    {
        double u[28];
        ev_c6_unique(load, u);
#pragma unroll
        for (int q = 0; q < 28; ++q) s.a6[q][threadIdx.x] = u[q];
    }
    if (WITH8) {
        double u[45];
        ev_c8_unique(load, u);
#pragma unroll
        for (int q = 0; q < 45; ++q) s.a8[q][threadIdx.x] = u[q];
    }
}

__device__ __forceinline__ void load_full15(const double2* __restrict__ nlm, long long ld, long long p, double2& n00, double2 n2[5],
                                            double2 n4[9]) {
    n00 = nlm[p];
#pragma unroll
    for (int m = 0; m < 5; ++m) n2[m] = nlm[(long long)(1 + m) * ld + p];
#pragma unroll
    for (int m = 0; m < 9; ++m) n4[m] = nlm[(long long)(6 + m) * ld + p];
}

// ---- a6 (N,3,3,3,3,3,3), Fortran order: plane index sum_d i_d 3^d
__global__ void __launch_bounds__(kB) a6_kernel(const double2* __restrict__ nlm, long long N, long long ld, double* __restrict__ out,
                                                long long ldo) {
    __shared__ Smem s;
    const long long p = (long long)blockIdx.x * kB + threadIdx.x;
    if (p >= N) return;
    stage_hi<false>(s, nlm, ld, p);
    for (int e = 0; e < 729; ++e) {
        int n0 = 0, n2 = 0, r = e;
#pragma unroll
        for (int d = 0; d < 6; ++d) { const int i = r % 3; r /= 3; n0 += (i == 0); n2 += (i == 2); }
        out[(long long)e * ldo + p] = s.a6[sfb::sym_index<6>(n0, n2)][threadIdx.x];
    }
}

// ---- Eij_tranisotropic with n' = 3: same contract as eij_kernel (sfb_fields.cu)
__global__ void __launch_bounds__(kB) eij3_kernel(const double2* __restrict__ nlm, long long N, long long ld,
                                                  const double* __restrict__ e1, const double* __restrict__ e2,
                                                  const double* __restrict__ e3, long long lde, sfb::EijCoef K,
                                                  double* __restrict__ Eij, long long ldo, double* __restrict__ ei_out,
                                                  double* __restrict__ lam_out, int* __restrict__ status) {
    __shared__ Smem s;
    const long long p = (long long)blockIdx.x * kB + threadIdx.x;
    if (p >= N) return;
    double2 n00, n2f[5], n4f[9];
    load_full15(nlm, ld, p, n00, n2f, n4f);
    double a4u[15];
    sfb::ev_c4_unique(n00, n2f, n4f, a4u);
    double2 n2[3], n4[5];
#pragma unroll
    for (int m = 0; m < 3; ++m) n2[m] = n2f[2 + m];
#pragma unroll
    for (int m = 0; m < 5; ++m) n4[m] = n4f[4 + m];
    double a2m[3][3];
    sfb::a2_from(n00, n2, a2m);
    stage_hi<true>(s, nlm, ld, p);

    double e[3][3];
    if (e1) {
#pragma unroll
        for (int x = 0; x < 3; ++x) { e[0][x] = e1[(long long)x * lde + p]; e[1][x] = e2[(long long)x * lde + p]; e[2][x] = e3[(long long)x * lde + p]; }
    } else {
        double lam[3];
        sfb::eigframe(a2m, 0, e, lam);
        if (ei_out) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                lam_out[(long long)i * ldo + p] = lam[i];
#pragma unroll
                for (int x = 0; x < 3; ++x) ei_out[(long long)(i + 3 * x) * ldo + p] = e[i][x];
            }
        }
    }
    const sfb::SymView A6{&s.a6[0][threadIdx.x], kB}, A8{&s.a8[0][threadIdx.x], kB};
    const sfb::RegView I6{g_iso.a6}, I8{g_iso.a8};
    double a2iso[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) a2iso[i][j] = i == j ? 1.0 / 3.0 : 0.0;       // f_ev_c2 of the isotropic state: exact
    double Es[6];
#pragma unroll 1
    for (int q = 0; q < 6; ++q) {
        const int iv = (q < 3) ? q : (q == 3 ? 1 : 0);
        const int iw = (q < 3) ? q : (q == 5 ? 1 : 2);
        double tau[3][3], vw[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double ev_i = 0, ev_j = 0, ew_i = 0, ew_j = 0;      // dynamic row select without local-memory indexing
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (r == iv) { ev_i = e[r][i]; ev_j = e[r][j]; }
                    if (r == iw) { ew_i = e[r][i]; ew_j = e[r][j]; }
                }
                vw[i][j] = ev_i * ew_j;
                tau[i][j] = q < 3 ? ((i == j) ? 1.0 / 3.0 : 0.0) - ev_i * ev_j : ev_i * ew_j + ew_i * ev_j;
            }
        double eps[3][3], epsi[3][3];
        sfb::sachs_n3_eps(tau, a2m, a4u, A6, A8, K.sA, K.sB, K.sC, eps);
        sfb::sachs_n3_eps(tau, a2iso, g_iso.a4, I6, I8, K.sA, K.sB, K.sC, epsi);
        const double es = sfb::dinner22(eps, vw) / sfb::dinner22(epsi, vw);
#pragma unroll
        for (int r = 0; r < 6; ++r)
            if (r == q) Es[r] = es;
    }
    double E[6];
    const int st = sfb::eij_tranisotropic<true>(n00, n2, n4, e, K, E, Es);
#pragma unroll
    for (int q = 0; q < 6; ++q) Eij[(long long)q * ldo + p] = E[q];
    if (status) status[p] = st;
}

// ---- E_CAFFE (Placidi et al. 2010): eps (N,3,3) Fortran order
__global__ void __launch_bounds__(kB) caffe_kernel(const double2* __restrict__ nlm, long long N, long long ld,
                                                   const double* __restrict__ eps, long long lde, double Emin, double Emax, int n_RSS,
                                                   double* __restrict__ E) {
    __shared__ Smem s;
    const long long p = (long long)blockIdx.x * kB + threadIdx.x;
    if (p >= N) return;
    double tau[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tau[i][j] = eps[(long long)(i + 3 * j) * lde + p];
    double D;
    if (n_RSS == 4) {
        double2 n00, n2f[5], n4f[9];
        load_full15(nlm, ld, p, n00, n2f, n4f);
        double a4u[15];
        sfb::ev_c4_unique(n00, n2f, n4f, a4u);
        stage_hi<true>(s, nlm, ld, p);
        const sfb::SymView A6{&s.a6[0][threadIdx.x], kB}, A8{&s.a8[0][threadIdx.x], kB};
        D = sfb::ev_D4(tau, a4u, A6, A8);
    } else {
        double2 n00, n2[3], n4[5];
        sfb::load_m_ge0(nlm, ld, p, n00, n2, n4);
        double tsq[3][3], tauv[6], tsqv[6];
        sfb::matmul3(tau, tau, tsq);
        sfb::mat_to_vec(tau, tauv);
        sfb::mat_to_vec(tsq, tsqv);
        D = sfb::ev_D2(n00, n2, n4, tauv, tsqv, tsq[0][0] + tsq[1][1] + tsq[2][2]);
    }
    const double Dmax = n_RSS == 4 ? 35 / 8.0 : 5 / 2.0;             // src/enhancementfactors.f90:312
    const double ex = 4.0 / n_RSS;
    const double Dmaxpow = pow(Dmax, ex);
    double r;
    if (D < 1) {
        const double gam = ex / Dmaxpow * (Emax - 1) / (1 - Emin);
        r = Emin + (1 - Emin) * pow(D, gam);
    } else {
        r = ((Emax - 1) * pow(D, ex) + Dmaxpow - Emax) / (Dmaxpow - 1);
    }
    E[p] = r;
}

// ---- pfJ = 4 Pi sum_l (2l+1) S(l),  S(l) = sum_m |n_l^m|^2 / (2l+1)
__global__ void __launch_bounds__(128) pfj_kernel(const double2* __restrict__ nlm, long long N, long long ld, int Lmax,
                                                  double* __restrict__ J) {
    const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
    if (p >= N) return;
    double acc = 0.0;
    for (int l = 0; l <= Lmax; l += 2) {
        const int j0 = l * (l + 1) / 2 - l;
        double sum = 0.0;
        for (int j = j0; j < j0 + 2 * l + 1; ++j) {
            const double2 v = nlm[(long long)j * ld + p];
            sum += v.x * v.x + v.y * v.y;
        }
        acc += (2 * l + 1) * (1.0 / (2 * l + 1) * sum);
    }
    J[p] = 4 * 3.141592653589793 * acc;
}

// ---- state ingest a2 / a4 / a6 -> nlm(1:6 / 1:15 / 1:28)  (src/moments.f90:68-92): affine map of the tensor entries,
// A (N, 3^k) Fortran order, table rows sorted so each output row is one running sum
template <int TAG>
__global__ void __launch_bounds__(128) ingest_kernel(const double* __restrict__ A, long long N, long long ld,
                                                     double2* __restrict__ out, long long ldo) {
    // the tables are __device__ symbols: they must be named in device code (a host-side address would be the shadow)
    const short* row = TAG == 0 ? kIngRow_0 : (TAG == 1 ? kIngRow_1 : kIngRow_2);
    const short* flat = TAG == 0 ? kIngFlat_0 : (TAG == 1 ? kIngFlat_1 : kIngFlat_2);
    const double* cr = TAG == 0 ? kIngRe_0 : (TAG == 1 ? kIngRe_1 : kIngRe_2);
    const double* ci = TAG == 0 ? kIngIm_0 : (TAG == 1 ? kIngIm_1 : kIngIm_2);
    const double* c0r = TAG == 0 ? kIngC0Re_0 : (TAG == 1 ? kIngC0Re_1 : kIngC0Re_2);
    const double* c0i = TAG == 0 ? kIngC0Im_0 : (TAG == 1 ? kIngC0Im_1 : kIngC0Im_2);
    constexpr int nnz = TAG == 0 ? SFB_ING_NNZ_0 : (TAG == 1 ? SFB_ING_NNZ_1 : SFB_ING_NNZ_2);
    constexpr int nrow = TAG == 0 ? SFB_ING_NROW_0 : (TAG == 1 ? SFB_ING_NROW_1 : SFB_ING_NROW_2);
    const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
    if (p >= N) return;
    int t = 0;
    for (int r = 0; r < nrow; ++r) {
        double2 acc = make_double2(c0r[r], c0i[r]);
        while (t < nnz && row[t] == r) {
            const double a = A[(long long)flat[t] * ld + p];
            acc.x = fma(cr[t], a, acc.x);
            acc.y = fma(ci[t], a, acc.y);
            ++t;
        }
        out[(long long)r * ldo + p] = acc;
    }
}

cudaError_t ensure_iso(cudaStream_t st) {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && done[dev]) return cudaSuccess;
    iso_kernel<<<1, 1, 0, st>>>();
    e = cudaGetLastError();
    if (e == cudaSuccess && dev < 64) done[dev] = true;
    return e;
}

inline unsigned nblk(long long N, int b) { return (unsigned)((N + b - 1) / b); }

}  // namespace

cudaError_t sfb_launch_a6(const double2* nlm, long long N, long long ld, double* out, long long ldo, cudaStream_t st) {
    if (N > 0) a6_kernel<<<nblk(N, kB), kB, 0, st>>>(nlm, N, ld, out, ldo);
    return cudaGetLastError();
}
cudaError_t sfb_launch_eij3(const double2* nlm, long long N, long long ld, const double* e1, const double* e2, const double* e3,
                            long long lde, const sfb::EijCoef& K, double* Eij, long long ldo, double* ei_out, double* lam_out,
                            int* status, cudaStream_t st) {
    cudaError_t e = ensure_iso(st);
    if (e != cudaSuccess) return e;
    if (N > 0) eij3_kernel<<<nblk(N, kB), kB, 0, st>>>(nlm, N, ld, e1, e2, e3, lde, K, Eij, ldo, ei_out, lam_out, status);
    return cudaGetLastError();
}
cudaError_t sfb_launch_caffe(const double2* nlm, long long N, long long ld, const double* eps, long long lde, double Emin, double Emax,
                             int n_grain, double* E, cudaStream_t st) {
    // src/enhancementfactors.f90:316-320: n' = 3 depends on RSS^4, anything else is treated as linear (RSS^2)
    if (N > 0) caffe_kernel<<<nblk(N, kB), kB, 0, st>>>(nlm, N, ld, eps, lde, Emin, Emax, n_grain == 3 ? 4 : 2, E);
    return cudaGetLastError();
}
cudaError_t sfb_launch_pfj(const double2* nlm, long long N, long long ld, int Lmax, double* J, cudaStream_t st) {
    if (N > 0) pfj_kernel<<<nblk(N, 128), 128, 0, st>>>(nlm, N, ld, Lmax, J);
    return cudaGetLastError();
}
cudaError_t sfb_launch_ingest(int rank, const double* A, long long N, long long ld, double2* out, long long ldo, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const unsigned nb = nblk(N, 128);
    if (rank == 2) ingest_kernel<0><<<nb, 128, 0, st>>>(A, N, ld, out, ldo);
    else if (rank == 4) ingest_kernel<1><<<nb, 128, 0, st>>>(A, N, ld, out, ldo);
    else ingest_kernel<2><<<nb, 128, 0, st>>>(A, N, ld, out, ldo);
    return cudaGetLastError();
}
