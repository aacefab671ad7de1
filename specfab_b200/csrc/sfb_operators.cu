// Batched operator export: M_LROT, M_DDRX_src, M_DDRX, M_REG as dense (N, n, n) arrays, the form the
// reference's Eulerian callers consume (src/specfabpy/fenics/CPO.py:200-202, firedrake/ice.py:202-204;
// SURVEY.md section 8f-1).  One thread per node, node-contiguous (coalesced) stores of every non-zero
// M(i,j); the per-L coefficient lists are merged on the host at init from the embedded Gaunt data with
// the algebra of codegen/operators.py:
//     M_LROT(i,j) = qe[D]*A(i,j) + (i*qo[D])*B(i,j),  D = m_i - m_j        src/dynamics.f90:78-96
//     M_DDRX_src(i,j) = sum_lk g[(lk,D)] * GC(i,j,(lk,D))                   src/dynamics.f90:295-297
#include <cmath>
#include <cstring>
#include <vector>

#include "sfb_common.cuh"
#include "sfb_moments.cuh"
#include "gen/tables.inc"

struct SfbNz {           // one structurally non-zero M(i,j): value = sum_t c[t] * f[fidx[t]]
    int i, j, nt;
    int fidx[3];
    double c[3];
};

namespace {

constexpr int kNMax = 231, kNCat = 15;
std::vector<double> g_dense[4];    // GC, GCm, GC_m1, GC_p1 dense [231][231][15] (host, built once)

void build_dense() {
    if (!g_dense[0].empty()) return;
    const int N[4] = {kGauntN_0, kGauntN_1, kGauntN_2, kGauntN_3};
    const unsigned short* I[4] = {kGauntI_0, kGauntI_1, kGauntI_2, kGauntI_3};
    const unsigned short* J[4] = {kGauntJ_0, kGauntJ_1, kGauntJ_2, kGauntJ_3};
    const unsigned char* K[4] = {kGauntK_0, kGauntK_1, kGauntK_2, kGauntK_3};
    const float* V[4] = {kGauntV_0, kGauntV_1, kGauntV_2, kGauntV_3};
    for (int t = 0; t < 4; ++t) {
        g_dense[t].assign((size_t)kNMax * kNMax * kNCat, 0.0);
        for (int e = 0; e < N[t]; ++e) g_dense[t][((size_t)I[t][e] * kNMax + J[t][e]) * kNCat + K[t][e]] = (double)V[t][e];
    }
}
inline double T(int t, int i, int j, int k) { return g_dense[t][((size_t)i * kNMax + j) * kNCat + k]; }

// one structurally non-zero entry (ii,jj) of the reduced operators (src/reducedform.f90:76-120): built from the
// full-form entries M(I_all(ii), +m_jj) [zp] and M(I_all(ii), -m_jj) [zn] (index into the SfbNz list, -1 = zero)
struct SfbRz {
    int ii, jj, s, zp, zn;
};

// the same, flattened for the export kernel: staged per CTA into shared memory with coalesced copies and read back as
// broadcasts (the warp-uniform GLOBAL table loads were what the kernel waited on)
struct SfbRzF {
    int o;                 // ii + r * jj
    int s;                 // sign of the (l, -m_jj) contribution: 0 (m_jj = 0), +1, -1
    int ntp, ntn;          // number of terms of M(I(ii), +m_jj) and M(I(ii), -m_jj); both 0 = structural zero
    int diag;              // ii == jj: M_DDRX subtracts <D> here
    int fp[3], fn[3];
    int pad;
    double cp[3], cn[3];
};
static_assert(sizeof(SfbRzF) == 96, "flat record: 6 x 16 bytes");
// dense (N, n, n) export: one record per output entry M(i, j), structural zeros included (nt = 0)
struct SfbNzF {
    int o;                 // i + n * j
    int nt, diag, pad;
    int fidx[3], pad2;
    double c[3], pad3;
};
static_assert(sizeof(SfbNzF) == 64, "flat record: 4 x 16 bytes");

struct DevLists {
    int L = 0;
    SfbNz *lrot = nullptr, *ddrx = nullptr;
    SfbRz *rlrot = nullptr, *rddrx = nullptr;
    SfbRzF *flrot = nullptr, *fddrx = nullptr;
    SfbNzF *dlrot = nullptr, *dddrx = nullptr;        // n^2 records each
    int n_lrot = 0, n_ddrx = 0, n_rlrot = 0, n_rddrx = 0;
} g_lists[64];

void reduced_list(int L, const std::vector<SfbNz>& nz, std::vector<SfbRz>& out) {
    const int n = (L + 1) * (L + 2) / 2;
    std::vector<int> where((size_t)n * n, -1);
    for (size_t z = 0; z < nz.size(); ++z) where[(size_t)nz[z].i * n + nz[z].j] = (int)z;
    int ii = 0;
    for (int li = 0; li <= L; li += 2)
        for (int mi = 0; mi <= li; ++mi, ++ii) {
            const int ip = li * (li + 1) / 2 + mi;
            int jj = 0;
            for (int lj = 0; lj <= L; lj += 2)
                for (int mj = 0; mj <= lj; ++mj, ++jj) {
                    const int jp = lj * (lj + 1) / 2 + mj, jn = jp - 2 * mj;
                    // every (ii, jj) is listed, structural zeros too (zp = zn = -1): the export kernel writes each output entry
                    // exactly once and needs no memset pass over the 4 r^2 8-byte planes per node
                    SfbRz r{ii, jj, mj == 0 ? 0 : ((mj & 1) ? -1 : 1), where[(size_t)ip * n + jp], where[(size_t)ip * n + jn]};
                    out.push_back(r);
                }
        }
}

// real(4) constants of src/dynamics.f90:79-86 promoted to double
const double SQRT3_F = 0x1.bb67aep+0, S56 = 0x1.d363d2p-1, S23 = 0x1.a20bd8p-1, S32 = 0x1.3988e2p+0;
const double C6 = 6.0 / 0x1.3988e2p+1;

void merge_lists(int L, std::vector<SfbNz>& lrot, std::vector<SfbNz>& ddrx) {
    build_dense();
    const int n = (L + 1) * (L + 2) / 2;
    std::vector<int> mm(n);
    {
        int j = 0;
        for (int l = 0; l <= L; l += 2)
            for (int m = -l; m <= l; ++m) mm[j++] = m;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int D = mm[i] - mm[j];
            if (D >= -2 && D <= 2) {
                double A = 0, B = 0;
                // same expressions as codegen/operators.py (GC=0, GCm=1, GC_m1=2, GC_p1=3)
                if (D == -2) A = -(3 * T(0, i, j, 1) - T(1, i, j, 1) + T(2, i, j, 2));
                if (D == -1) A = -(3 * T(0, i, j, 2) + S56 * T(2, i, j, 0) + S23 * T(2, i, j, 3) + 2 * T(3, i, j, 1));
                if (D == 0) A = -(3 * T(0, i, j, 3) + S32 * T(2, i, j, 4) + S32 * T(3, i, j, 2));
                if (D == 1) A = -(3 * T(0, i, j, 4) + 2 * T(2, i, j, 5) + S56 * T(3, i, j, 0) + S23 * T(3, i, j, 3));
                if (D == 2) A = -(3 * T(0, i, j, 5) + T(1, i, j, 5) + T(3, i, j, 4));
                if (D == 0) B = SQRT3_F * T(1, i, j, 0);
                if (D == -1) B = C6 * T(2, i, j, 0);
                if (D == 1) B = -C6 * T(3, i, j, 0);
                if (A != 0 || B != 0) {
                    SfbNz z{}; z.i = i; z.j = j; z.nt = 0;
                    if (A != 0) { z.fidx[z.nt] = D + 2; z.c[z.nt++] = A; }
                    if (B != 0) { z.fidx[z.nt] = 5 + D + 1; z.c[z.nt++] = B; }
                    lrot.push_back(z);
                }
            }
            if (D >= -4 && D <= 4) {
                SfbNz z{}; z.i = i; z.j = j; z.nt = 0;
                const int lks[3] = {0, 2, 4}, base[3] = {0, 3, 10};     // k(lk,mk) = base + mk
                for (int q = 0; q < 3; ++q) {
                    if (lks[q] < (D < 0 ? -D : D)) continue;
                    const int k = base[q] + D;
                    const double c = T(0, i, j, k);
                    if (c != 0) { z.fidx[z.nt] = 8 + k; z.c[z.nt++] = c; }
                }
                if (z.nt) ddrx.push_back(z);
            }
        }
}

#define SFB_L 20          // sfb_step_common.cuh needs compile-time names; only the forcing helpers are used here
#define SFB_DDRX 1
constexpr int kTN = 128;   // nodes per block
constexpr int kNF = 23;
constexpr int kThreads = 128;
__constant__ SfbRegConst c_reg;
enum { SC_C0 = 0, SC_LAM = 1, SC_RM = 2, SC_G0 = 3, SC_TAUV = 4, SC_TSQV = 10, SC_NORM = 16 };
#include "sfb_step_common.cuh"

// mode 0: M_LROT(eps, omg, iota, zeta)   mode 1: M_DDRX_src(tau)   mode 2: M_DDRX(nlm, tau)
// forcing coefficients of node p into f[.][t]; returns <D> (mode 2)
__device__ __forceinline__ double mexport_forcing(int mode, double2 (*f)[kTN], int t, long long p,
                                                  const double* __restrict__ a33, const double* __restrict__ b33,
                                                  const double2* __restrict__ nlm, long long ldn, long long ld, double iota, double zeta) {
    double davg = 0.0;
    {
        if (mode == 0) {
            // reference reads eps / omg entries directly: quad_rr(iota*eps + zetanorm*eps^2), quad_tp(omg)
            double e[3][3], w[3][3], sq[3][3], E[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) { e[i][j] = a33[(long long)(i + 3 * j) * ld + p]; w[i][j] = b33[(long long)(i + 3 * j) * ld + p]; }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) sq[i][j] = e[i][0] * e[0][j] + e[i][1] * e[1][j] + e[i][2] * e[2][j];
            const double zetanorm = zeta / sqrt(sq[0][0] + sq[1][1] + sq[2][2]);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) E[i][j] = iota * e[i][j] + zetanorm * sq[i][j];
            double2 qe[5], qo[3];
            quad_rr(E, qe);
            quad_tp(w, qo);
#pragma unroll
            for (int d = 0; d < 5; ++d) f[d][t] = qe[d];
#pragma unroll
            for (int d = 0; d < 3; ++d) f[5 + d][t] = make_double2(-qo[d].y, qo[d].x);     // i*qo
        } else {
            double T3[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) T3[i][j] = a33[(long long)(i + 3 * j) * ld + p];
            double2 qt[5], g[15];
            quad_rr(T3, qt);
            ddrx_weights_raw(qt, g);
            double dd = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) dd += T3[i][j] * T3[j][i];
            const double kk = 0x1.14d2dcd9ceb17p-3;
#pragma unroll
            for (int k = 0; k < 15; ++k) f[8 + k][t] = make_double2(((kk * g[k].x) * 5) / dd, ((kk * g[k].y) * 5) / dd);
            if (mode == 2) {      // <D>  (src/dynamics.f90:402-422)
                double sq[3][3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) sq[i][j] = T3[i][0] * T3[0][j] + T3[i][1] * T3[1][j] + T3[i][2] * T3[2][j];
                const double tv[6] = {T3[0][0], T3[1][1], T3[2][2], SFB_SQRT2 * T3[1][2], SFB_SQRT2 * T3[0][2], SFB_SQRT2 * T3[0][1]};
                const double sv[6] = {sq[0][0], sq[1][1], sq[2][2], SFB_SQRT2 * sq[1][2], SFB_SQRT2 * sq[0][2], SFB_SQRT2 * sq[0][1]};
                double2 n2[3], n4[5];
#pragma unroll
                for (int m = 0; m < 3; ++m) n2[m] = nlm[(long long)(3 + m) * ldn + p];
#pragma unroll
                for (int m = 0; m < 5; ++m) n4[m] = nlm[(long long)(10 + m) * ldn + p];
                davg = sfb::ev_D2(nlm[p], n2, n4, tv, sv, sq[0][0] + sq[1][1] + sq[2][2]);
            }
        }
    }
    return davg;
}

__device__ __forceinline__ double2 nz_value(const SfbNz& e, const double2 (*f)[kTN], int t, int mode, double davg) {
    double2 v = make_double2(0.0, 0.0);
    for (int q = 0; q < e.nt; ++q) {
        const double2 ff = f[e.fidx[q]][t];
        v.x = fma(e.c[q], ff.x, v.x);
        v.y = fma(e.c[q], ff.y, v.y);
    }
    if (mode == 2 && e.i == e.j) v.x -= davg;
    return v;
}

constexpr int kNzBatch = 160;     // flat records staged per pass (10 KB of shared memory)
__global__ void __launch_bounds__(kThreads) mexport_kernel(int mode, const SfbNzF* __restrict__ nf, int nent,
                                                           const double* __restrict__ a33, const double* __restrict__ b33,
                                                           const double2* __restrict__ nlm, long long ldn, long long N, long long ld,
                                                           double iota, double zeta, double2* __restrict__ M, long long ldm) {
    // same scheme as mexport_reduced_kernel below: every one of the n^2 entries is written exactly once (no memset pass), the
    // CTA's contiguous share of the table is staged through shared memory, the forcing preparation is amortised over it
    __shared__ double2 fbuf[15][kTN];
    __shared__ __align__(16) SfbNzF recs[kNzBatch];
    double2 (*f)[kTN] = fbuf - (mode == 0 ? 0 : 8);
    const int t = threadIdx.x;
    const long long p = (long long)blockIdx.x * kTN + t;
    const bool valid = p < N;
    const double davg = valid ? mexport_forcing(mode, f, t, p, a33, b33, nlm, ldn, ld, iota, zeta) : 0.0;
    const int per = (nent + gridDim.y - 1) / gridDim.y, z0 = blockIdx.y * per, z1 = min(nent, z0 + per);
    for (int zb = z0; zb < z1; zb += kNzBatch) {
        const int nb = min(kNzBatch, z1 - zb);
        __syncthreads();
        {
            const int4* src = reinterpret_cast<const int4*>(nf + zb);
            int4* dst = reinterpret_cast<int4*>(recs);
            for (int q = t; q < nb * 4; q += kThreads) dst[q] = src[q];
        }
        __syncthreads();
        if (!valid) continue;
        for (int z = 0; z < nb; ++z) {
            const SfbNzF& e = recs[z];
            double2 v = make_double2(0.0, 0.0);
            for (int q = 0; q < e.nt; ++q) { const double2 ff = f[e.fidx[q]][t]; v.x = fma(e.c[q], ff.x, v.x); v.y = fma(e.c[q], ff.y, v.y); }
            if (mode == 2 && e.diag && e.nt) v.x -= davg;
            __stcs(M + (long long)e.o * ldm + p, v);
        }
    }
}

// the same operators written directly in reduced form (Mrr, Mri, Mir, Mii each (N, r, r) real), src/reducedform.f90:76-120
constexpr int kRzBatch = 96;      // flat records staged per pass (9 KB of shared memory)
__global__ void __launch_bounds__(kThreads) mexport_reduced_kernel(int mode, const SfbRzF* __restrict__ rf, int nrz, int r,
                                                                   const double* __restrict__ a33,
                                                                   const double* __restrict__ b33, const double2* __restrict__ nlm,
                                                                   long long ldn, long long N, long long ld, double iota, double zeta,
                                                                   double* __restrict__ Mrr, double* __restrict__ Mri,
                                                                   double* __restrict__ Mir, double* __restrict__ Mii, long long ldm) {
    // forcing entries 0..7 (M_LROT) or 8..22 (M_DDRX*): at most 15 of the kNF rows are live in one launch
    __shared__ double2 fbuf[15][kTN];
    __shared__ __align__(16) SfbRzF recs[kRzBatch];
    double2 (*f)[kTN] = fbuf - (mode == 0 ? 0 : 8);
    const int t = threadIdx.x;
    const long long p = (long long)blockIdx.x * kTN + t;
    const bool valid = p < N;
    const double davg = valid ? mexport_forcing(mode, f, t, p, a33, b33, nlm, ldn, ld, iota, zeta) : 0.0;
    // the forcing preparation above is repeated by every blockIdx.y: each CTA takes a contiguous share of the r^2 entries of
    // the dense list and writes every one of them exactly once, structural zeros included (no memset pass over the output)
    const int per = (nrz + gridDim.y - 1) / gridDim.y, z0 = blockIdx.y * per, z1 = min(nrz, z0 + per);
    for (int zb = z0; zb < z1; zb += kRzBatch) {
        const int nb = min(kRzBatch, z1 - zb);
        __syncthreads();
        {   // coalesced 16-byte copies of nb records
            const int4* src = reinterpret_cast<const int4*>(rf + zb);
            int4* dst = reinterpret_cast<int4*>(recs);
            for (int q = t; q < nb * 6; q += kThreads) dst[q] = src[q];
        }
        __syncthreads();
        if (!valid) continue;
        for (int z = 0; z < nb; ++z) {
            const SfbRzF& e = recs[z];
            const long long o = (long long)e.o * ldm + p;
            if ((e.ntp | e.ntn) == 0) {
                __stcs(Mrr + o, 0.0); __stcs(Mri + o, 0.0); __stcs(Mir + o, 0.0); __stcs(Mii + o, 0.0);
                continue;
            }
            double2 vp = make_double2(0.0, 0.0), vn = make_double2(0.0, 0.0);
            for (int q = 0; q < e.ntp; ++q) { const double2 ff = f[e.fp[q]][t]; vp.x = fma(e.cp[q], ff.x, vp.x); vp.y = fma(e.cp[q], ff.y, vp.y); }
            for (int q = 0; q < e.ntn; ++q) { const double2 ff = f[e.fn[q]][t]; vn.x = fma(e.cn[q], ff.x, vn.x); vn.y = fma(e.cn[q], ff.y, vn.y); }
            if (mode == 2 && e.diag) vp.x -= davg;
            const double s = (double)e.s;
            __stcs(Mrr + o, vp.x + s * vn.x);
            __stcs(Mri + o, -vp.y + s * vn.y);
            __stcs(Mir + o, vp.y + s * vn.y);
            __stcs(Mii + o, vp.x - s * vn.x);
        }
    }
}

// reduce_M on an existing dense operator M (N, n, n), complex or real; blockIdx.y = reduced entry (ii + r*jj)
__global__ void __launch_bounds__(kThreads) reduce_dense_kernel(const double* __restrict__ M, int is_complex, int L, long long N,
                                                                long long ldi, double* __restrict__ Mrr, double* __restrict__ Mri,
                                                                double* __restrict__ Mir, double* __restrict__ Mii, long long ldo) {
    const long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (p >= N) return;
    const int n = (L + 1) * (L + 2) / 2, r = (L + 2) * (L + 2) / 4;
    for (int z = blockIdx.y; z < r * r; z += gridDim.y) {
        const int ii = z % r, jj = z / r;
        int li = 0, lj = 0;
        while ((li / 2 + 1) * (li / 2 + 1) <= ii) li += 2;          // reduced index of (l, m) is (l/2)^2 + m
        while ((lj / 2 + 1) * (lj / 2 + 1) <= jj) lj += 2;
        const int mi = ii - (li / 2) * (li / 2), mj = jj - (lj / 2) * (lj / 2);
        const int ip = li * (li + 1) / 2 + mi, jp = lj * (lj + 1) / 2 + mj, jn = jp - 2 * mj;
        double2 vp, vn;
        if (is_complex) {
            const double2* Mc = reinterpret_cast<const double2*>(M);
            vp = Mc[((long long)ip + (long long)n * jp) * ldi + p];
            vn = Mc[((long long)ip + (long long)n * jn) * ldi + p];
        } else {
            vp = make_double2(M[((long long)ip + (long long)n * jp) * ldi + p], 0.0);
            vn = make_double2(M[((long long)ip + (long long)n * jn) * ldi + p], 0.0);
        }
        const double s = mj == 0 ? 0.0 : ((mj & 1) ? -1.0 : 1.0);
        const long long o = (long long)z * ldo + p;
        Mrr[o] = vp.x + s * vn.x;
        Mri[o] = -vp.y + s * vn.y;
        Mir[o] = vp.y + s * vn.y;
        Mii[o] = vp.x - s * vn.x;
    }
}

__global__ void mreg_kernel(const double* __restrict__ eps, long long N, long long ld, int n, int L, double nu,
                            const double* __restrict__ regdiag_l, double* __restrict__ M, long long ldm) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    double fro = 0.0;
    for (int q = 0; q < 9; ++q) { const double v = eps[(long long)q * ld + p]; fro += v * v; }
    const double ratemag = nu * sqrt(fro);                      // src/dynamics.f90:516
    int j = 0;
    for (int l = 0; l <= L; l += 2)
        for (int m = -l; m <= l; ++m, ++j) M[((long long)j + (long long)n * j) * ldm + p] = -ratemag * regdiag_l[l / 2];
}

}  // namespace

cudaError_t sfb_ops_prepare(int L) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    DevLists& d = g_lists[dev];
    if (d.L == L) return cudaSuccess;
    std::vector<SfbNz> a, b;
    merge_lists(L, a, b);
    std::vector<SfbRz> ra, rb;
    reduced_list(L, a, ra);
    reduced_list(L, b, rb);
    const int r = (L + 2) * (L + 2) / 4;
    auto flatten = [&](const std::vector<SfbNz>& nz, const std::vector<SfbRz>& rz) {
        std::vector<SfbRzF> out(rz.size());
        for (size_t z = 0; z < rz.size(); ++z) {
            SfbRzF f{};
            f.o = rz[z].ii + r * rz[z].jj; f.s = rz[z].s; f.diag = rz[z].ii == rz[z].jj;
            if (rz[z].zp >= 0) { const SfbNz& e = nz[rz[z].zp]; f.ntp = e.nt; for (int q = 0; q < e.nt; ++q) { f.fp[q] = e.fidx[q]; f.cp[q] = e.c[q]; } }
            if (rz[z].zn >= 0 && rz[z].s != 0) { const SfbNz& e = nz[rz[z].zn]; f.ntn = e.nt; for (int q = 0; q < e.nt; ++q) { f.fn[q] = e.fidx[q]; f.cn[q] = e.c[q]; } }
            out[z] = f;
        }
        return out;
    };
    const std::vector<SfbRzF> fa = flatten(a, ra), fb = flatten(b, rb);
    const int nn = (L + 1) * (L + 2) / 2;
    auto densify = [&](const std::vector<SfbNz>& nz) {
        std::vector<SfbNzF> out((size_t)nn * nn);
        for (int j = 0; j < nn; ++j)
            for (int i = 0; i < nn; ++i) { SfbNzF f{}; f.o = i + nn * j; f.diag = i == j; out[(size_t)i + (size_t)nn * j] = f; }
        for (const SfbNz& e : nz) {
            SfbNzF& f = out[(size_t)e.i + (size_t)nn * e.j];
            f.nt = e.nt;
            for (int q = 0; q < e.nt; ++q) { f.fidx[q] = e.fidx[q]; f.c[q] = e.c[q]; }
        }
        return out;
    };
    const std::vector<SfbNzF> da = densify(a), db = densify(b);
    cudaFree(d.lrot); cudaFree(d.ddrx); cudaFree(d.rlrot); cudaFree(d.rddrx); cudaFree(d.flrot); cudaFree(d.fddrx); cudaFree(d.dlrot); cudaFree(d.dddrx);
    d = DevLists();
    cudaError_t e;
    if ((e = cudaMalloc(&d.dlrot, da.size() * sizeof(SfbNzF))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.dddrx, db.size() * sizeof(SfbNzF))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.dlrot, da.data(), da.size() * sizeof(SfbNzF), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.dddrx, db.data(), db.size() * sizeof(SfbNzF), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.flrot, fa.size() * sizeof(SfbRzF))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.fddrx, fb.size() * sizeof(SfbRzF))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.flrot, fa.data(), fa.size() * sizeof(SfbRzF), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.fddrx, fb.data(), fb.size() * sizeof(SfbRzF), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.rlrot, ra.size() * sizeof(SfbRz))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.rddrx, rb.size() * sizeof(SfbRz))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.rlrot, ra.data(), ra.size() * sizeof(SfbRz), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.rddrx, rb.data(), rb.size() * sizeof(SfbRz), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    d.n_rlrot = (int)ra.size(); d.n_rddrx = (int)rb.size();
    if ((e = cudaMalloc(&d.lrot, a.size() * sizeof(SfbNz))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d.ddrx, b.size() * sizeof(SfbNz))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.lrot, a.data(), a.size() * sizeof(SfbNz), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d.ddrx, b.data(), b.size() * sizeof(SfbNz), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    d.n_lrot = (int)a.size(); d.n_ddrx = (int)b.size(); d.L = L;
    return cudaSuccess;
}

void sfb_ops_release() {
    for (auto& d : g_lists) { cudaFree(d.lrot); cudaFree(d.ddrx); cudaFree(d.rlrot); cudaFree(d.rddrx); cudaFree(d.flrot); cudaFree(d.fddrx); cudaFree(d.dlrot); cudaFree(d.dddrx); d = DevLists(); }
}

// M must be zero-filled by the caller side of the launcher (done here with cudaMemsetAsync)
cudaError_t sfb_launch_mexport(int mode, int L, const double* a33, const double* b33, const double2* nlm, long long ldn,
                               long long N, long long ld, double iota, double zeta, double2* M, cudaStream_t st) {
    cudaError_t e = sfb_ops_prepare(L);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    const DevLists& d = g_lists[dev & 63];
    const int n = (L + 1) * (L + 2) / 2;
    if (N <= 0) return cudaSuccess;
    const SfbNzF* nf = mode == 0 ? d.dlrot : d.dddrx;
    const int nent = n * n;
    const long long xt = (N + kTN - 1) / kTN;
    int gy = std::max(1, std::min((nent + 191) / 192, 64));
    while (gy < 64 && xt * gy < 4 * 148) ++gy;
    dim3 grid((unsigned)xt, (unsigned)gy);
    mexport_kernel<<<grid, kThreads, 0, st>>>(mode, nf, nent, a33, b33, nlm, ldn, N, ld, iota, zeta, M, N);
    return cudaGetLastError();
}

cudaError_t sfb_launch_mexport_reduced(int mode, int L, const double* a33, const double* b33, const double2* nlm, long long ldn,
                                       long long N, long long ld, double iota, double zeta, double* Mrr, double* Mri, double* Mir,
                                       double* Mii, cudaStream_t st) {
    cudaError_t e = sfb_ops_prepare(L);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    const DevLists& d = g_lists[dev & 63];
    const int r = (L + 2) * (L + 2) / 4;
    if (N <= 0) return cudaSuccess;
    const SfbRzF* rf = mode == 0 ? d.flrot : d.fddrx;
    const int nrz = mode == 0 ? d.n_rlrot : d.n_rddrx;
    // nrz = r^2 (dense list).  Entries per CTA: enough to amortise the per-node forcing preparation (M_DDRX: weights and <D>),
    // few enough to keep >= ~4 CTAs per SM in flight on small batches
    const int kRzPerCta = 96;
    const long long xt = (N + kTN - 1) / kTN;
    int gy = std::max(1, std::min((nrz + kRzPerCta - 1) / kRzPerCta, 64));
    while (gy < 64 && gy < nrz && xt * gy < 4 * 148) ++gy;
    dim3 grid((unsigned)xt, (unsigned)gy);
    mexport_reduced_kernel<<<grid, kThreads, 0, st>>>(mode, rf, nrz, r, a33, b33, nlm, ldn, N, ld, iota, zeta, Mrr, Mri, Mir, Mii, N);
    return cudaGetLastError();
}

cudaError_t sfb_launch_reduce_dense(const double* M, int is_complex, int L, long long N, long long ldi, double* Mrr, double* Mri,
                                    double* Mir, double* Mii, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    const int r = (L + 2) * (L + 2) / 4;
    dim3 grid((unsigned)((N + kThreads - 1) / kThreads), (unsigned)std::min(r * r, 256));
    reduce_dense_kernel<<<grid, kThreads, 0, st>>>(M, is_complex, L, N, ldi, Mrr, Mri, Mir, Mii, N);
    return cudaGetLastError();
}

cudaError_t sfb_launch_mreg(int L, const SfbRegConst& reg, const double* eps, long long N, long long ld, double* M, cudaStream_t st) {
    const int n = (L + 1) * (L + 2) / 2;
    cudaError_t e;
    if ((e = cudaMemsetAsync(M, 0, (size_t)N * n * n * sizeof(double), st)) != cudaSuccess) return e;
    if (N <= 0) return cudaSuccess;
    double* dreg = nullptr;
    if ((e = cudaMallocAsync(&dreg, sizeof(reg.regdiag), st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(dreg, reg.regdiag, sizeof(reg.regdiag), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    mreg_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(eps, N, ld, n, L, reg.nu, dreg, M, N);
    e = cudaGetLastError();
    cudaFreeAsync(dreg, st);
    return e;
}
