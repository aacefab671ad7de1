// Higher-order structure tensors a6, a8 and the closures built on them:
//   ev_D4            src/dynamics.f90:424-448            (basal-plane RSS^4, CAFFE with n'=3)
//   Sachs, n'=3      src/homogenizations.f90:93-102,115  (and :212-218 for the isotropic denominator)
// a2, a6 and a8 are exactly symmetric in the reference (tools/make_moment_tables.py measures the permutation
// deviation of its 3^k separate assignments: 0), so they are held as unique entries indexed by the counts
// (n0, n1, n2) of the index values; a4 is NOT symmetric there (alias quirk, src/include/ev_c4__body.f90:78) and
// is therefore always addressed through a4_at(l,k,i,j).
#pragma once
#include "sfb_fields.cuh"
#include "gen/moments_hi.inc"

namespace sfb {

// rank of the sorted index tuple with n0 zeros and n2 twos among the K-tuples (lexicographic)
template <int K>
__host__ __device__ constexpr int sym_index(int n0, int n2) { return (K - n0) * (K - n0 + 1) / 2 + n2; }
template <int K>
__host__ __device__ constexpr int sym_count() { return (K + 1) * (K + 2) / 2; }

// unique-entry view of one node's tensor: element u lives at p[u * stride]
struct SymView {
    const double* p;
    int stride;
    __device__ __forceinline__ double operator[](int u) const { return p[u * stride]; }
};
struct RegView {
    const double* p;
    __device__ __forceinline__ double operator[](int u) const { return p[u]; }
};

// out (rank K-2, unique) = in (rank K, unique) : B, i.e. out_{...} = sum_ij in_{...ij} B_ji   (doubleinner62/82)
template <int K, class In>
__device__ __forceinline__ void sym_contract(const In& in, const double B[3][3], double* out) {
#pragma unroll
    for (int n0 = K - 2; n0 >= 0; --n0)
#pragma unroll
        for (int n1 = K - 2 - n0; n1 >= 0; --n1) {
            const int n2 = K - 2 - n0 - n1;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    acc = fma(in[sym_index<K>(n0 + (i == 0) + (j == 0), n2 + (i == 2) + (j == 2))], B[j][i], acc);
            out[sym_index<K - 2>(n0, n2)] = acc;
        }
}

// unique rank-2 entries (00,01,02,11,12,22) -> full symmetric matrix
__device__ __forceinline__ void sym2_to_mat(const double u[6], double m[3][3]) {
    m[0][0] = u[0]; m[0][1] = m[1][0] = u[1]; m[0][2] = m[2][0] = u[2];
    m[1][1] = u[3]; m[1][2] = m[2][1] = u[4]; m[2][2] = u[5];
}

// a4(l,k,i,j) of the reference from its 15 unique values, alias quirk included
__device__ __forceinline__ double a4_at(const double u[15], int l, int k, int i, int j) {
    const int q = (l == 2 && k == 1 && i == 0 && j == 1) ? 8 : a4_unique_index(l, k, i, j);
    return u[q];
}
// doubleinner42(a4, B)(l,k) = sum_ij a4(l,k,i,j) B(j,i)        src/tensorproducts.f90:169-179
__device__ __forceinline__ void a4_contract(const double u[15], const double B[3][3], double out[3][3]) {
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc = fma(a4_at(u, l, k, i, j), B[j][i], acc);
            out[l][k] = acc;
        }
}

__device__ __forceinline__ void matmul3(const double A[3][3], const double B[3][3], double C[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}

// eps of rheo_fwd_tranisotropic_sachshomo for n' = 3 (src/homogenizations.f90:93-102,115).  a2m full matrix,
// a4u 15 unique (+quirk), A6 / A8 unique-entry views.  cA, cB, cC from rheo_params_tranisotropic(.., n=3, ef=+1).
template <class V6, class V8>
__device__ __forceinline__ void sachs_n3_eps(const double tau[3][3], const double a2m[3][3], const double a4u[15],
                                             const V6& A6, const V8& A8, double cA, double cB, double cC,
                                             double eps[3][3]) {
    double tausq[3][3];
    matmul3(tau, tau, tausq);
    const double I2 = dinner22(tau, tau);
    double c4t[3][3], c4q[3][3];
    a4_contract(a4u, tau, c4t);            // ev_c4 : tau
    a4_contract(a4u, tausq, c4q);          // ev_c4 : tausq
    // a6 : tau : tau,  a6 : tausq : tau
    double t4[15], r2[6], c6tt[3][3], c6qt[3][3];
    sym_contract<6>(A6, tau, t4);
    sym_contract<4>(RegView{t4}, tau, r2);
    sym2_to_mat(r2, c6tt);
    sym_contract<6>(A6, tausq, t4);
    sym_contract<4>(RegView{t4}, tau, r2);
    sym2_to_mat(r2, c6qt);
    // a8 : tau : tau : tau
    double t6[28], c8ttt[3][3];
    sym_contract<8>(A8, tau, t6);
    sym_contract<6>(RegView{t6}, tau, t4);
    sym_contract<4>(RegView{t4}, tau, r2);
    sym2_to_mat(r2, c8ttt);

    const double etac0 = I2 * 1.0 + cB * dinner22(c4t, tau) + 2 * cC * dinner22(a2m, tausq);
    double etac2[3][3], a4tau[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            etac2[i][j] = I2 * a2m[i][j] + cB * c6tt[i][j] + 2 * cC * c4q[i][j];
            a4tau[i][j] = I2 * c4t[i][j] + cB * c8ttt[i][j] + 2 * cC * c6qt[i][j];
        }
    const double e2t = dinner22(etac2, tau);
    double te[3][3], et[3][3];
    matmul3(tau, etac2, te);
    matmul3(etac2, tau, et);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            eps[i][j] = etac0 * tau[i][j] - cA * e2t * (i == j ? 1.0 : 0.0) + cB * a4tau[i][j] + cC * (te[i][j] + et[i][j]);
}

// ev_D4 (src/dynamics.f90:424-448): 35/2 [ tausq:(a4:tausq) + a8::::tau^4 - 2 (a6:tau:tau):tausq ] / (tau:tau)^2
template <class V6, class V8>
__device__ __forceinline__ double ev_D4(const double tau[3][3], const double a4u[15], const V6& A6, const V8& A8) {
    double tausq[3][3];
    matmul3(tau, tau, tausq);
    const double norm = tausq[0][0] + tausq[1][1] + tausq[2][2];
    double c4q[3][3];
    a4_contract(a4u, tausq, c4q);
    double D = dinner22(tausq, c4q);
    double t6[28], t4[15], r2[6], m[3][3];
    sym_contract<8>(A8, tau, t6);
    sym_contract<6>(RegView{t6}, tau, t4);
    sym_contract<4>(RegView{t4}, tau, r2);
    sym2_to_mat(r2, m);
    D = D + dinner22(m, tau);
    sym_contract<6>(A6, tau, t4);
    sym_contract<4>(RegView{t4}, tau, r2);
    sym2_to_mat(r2, m);
    D = D - 2 * dinner22(m, tausq);
    return 35 / 2.0 * D / (norm * norm);
}

}  // namespace sfb
