// Eij_orthotropic_arr: eigenenhancements of a polycrystal of orthotropic grains (olivine), Sachs homogenisation.
// Reference: src/specfabpy.f90:488-500 (array entry), src/enhancementfactors.f90:134-189 (Eij / Evw_orthotropic),
// src/moments.f90:260-311,357-384 (a4_orth, a4_joint, a4_jointcross, ai_orthotropic),
// src/homogenizations.f90:288-319 (rheo_fwd_orthotropic_sachshomo), src/tensorproducts.f90:53-82,169-179.
//
// The reference evaluates, per node and per Eij component, three generated bilinear forms
//     ev(e) = REAL( sum_t C_t b_{p_t} n_{q_t} ) * k / norm         (81 entries, 2.4k-5.6k terms each)
// Every C_t is purely real or purely imaginary, so with P = (Re, Im) of the 15 x 15 products b_p n_q of one node,
// ev(e) = sum_t coef_t * P[idx_t].  Work decomposition here: one CTA owns TN nodes; thread <-> tensor entry e
// (81 of 96 lanes, one entry order for all tables, sorted by row length so the three warps are balanced); P of the TN nodes lives in shared memory,
// the ELL table streams through L1/L2 once per CTA (term-major, coalesced) and each term feeds TN DFMAs.
// Everything downstream of the moment tensors is linear in them, so each lane only keeps three running
// combinations per node (the a4_ii part, the sym2 part and the sym4 part of Q) for numerator and isotropic
// denominator, weights them with vw (x) tau of the six (v,w) pairs and the CTA reduces over entries at the end.
#include <cmath>
#include <cstdlib>

#include "sfb_fields.cuh"
#include "gen/orth_tables.inc"

namespace {

constexpr int kLanes = SFB_ORTH_LANES;

struct OrthCoef {                        // host-computed, src/rheologies.f90:186-206 + homogenizations.f90:303-304
    double cA[3];                        // weight of a4_ii(j) in Q: sum of ci(i)/4 over the M2(i) that contain it
    double cJ2[3];                       // weight of a4_sym2(a4_jk(i)):  -ci(i)/2
    double cJ4[3];                       // weight of a4_sym4(a4_jk(i)):   ci(3+i)
    int linear;                          // n_grain == 1
};

struct TabRef {
    const double* coef; const unsigned short* idx; const unsigned char* perm; const unsigned short* len;
    const double* ncoef; const unsigned short* nidx; int nnorm; double k;
    const double* iso_coef; const unsigned char* iso_im;
};
__device__ __forceinline__ TabRef tab_c2b2() { return {kOrthCoef_0, kOrthIdx_0, kOrthPerm_0, kOrthLen_0, kOrthNormCoef_0, kOrthNormIdx_0, SFB_ORTH_NNORM_0, SFB_ORTH_K_0, kOrthIsoCoef_0, kOrthIsoIm_0}; }
__device__ __forceinline__ TabRef tab_v4()   { return {kOrthCoef_1, kOrthIdx_1, kOrthPerm_1, kOrthLen_1, kOrthNormCoef_1, kOrthNormIdx_1, SFB_ORTH_NNORM_1, SFB_ORTH_K_1, kOrthIsoCoef_1, kOrthIsoIm_1}; }
__device__ __forceinline__ TabRef tab_c2v2() { return {kOrthCoef_2, kOrthIdx_2, kOrthPerm_2, kOrthLen_2, kOrthNormCoef_2, kOrthNormIdx_2, SFB_ORTH_NNORM_2, SFB_ORTH_K_2, kOrthIsoCoef_2, kOrthIsoIm_2}; }

template <int TN>
struct Smem {
    double2 q[3][TN][15];
    double P[450][TN];          // (Re, Im) of b_p n_q, node-minor: one term reads the TN nodes with vector loads
    double U[TN][4][15];        // unique a4 entries of q1, q2, q3 and of the isotropic state
    double F[TN][6][18];        // per Eij component: vw(3,3) then tau(3,3), row-major
    double red[3][TN * 12];
    int has3[TN];               // REAL(qlm_3(1)) > 1e-8  (src/moments.f90:371)
    int iso3[TN];               // the same test on the isotropic stand-in, qlm_iso(1) = qlm_1(1)
};

// P <- products b_p n_q of pair (ib, in) for all TN nodes
template <int TN>
__device__ __forceinline__ void make_P(Smem<TN>& s, int ib, int in) {
    for (int w = threadIdx.x; w < TN * 225; w += kLanes) {
        const int pq = w / TN, i = w - pq * TN, p = pq / 15, q = pq - p * 15;
        const double2 b = s.q[ib][i][p], n = s.q[in][i][q];
        s.P[2 * pq][i] = b.x * n.x - b.y * n.y;
        s.P[2 * pq + 1][i] = b.x * n.y + b.y * n.x;
    }
}

// ev(e) * k / norm of one table for the TN nodes (lane's own entry)
template <int TN>
__device__ __forceinline__ void eval_table(const TabRef& T, const Smem<TN>& s, int lane, double ev[TN]) {
    const int len = T.len[lane];
    double acc[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) acc[i] = 0.0;
    const double* cp = T.coef + lane;
    const unsigned short* ip = T.idx + lane;
#pragma unroll 4
    for (int t = 0; t < len; ++t) {
        const double c = cp[t * kLanes];
        const double2* pr = reinterpret_cast<const double2*>(s.P[ip[t * kLanes]]);
#pragma unroll
        for (int i = 0; i < TN / 2; ++i) {
            const double2 v = pr[i];
            acc[2 * i] = fma(c, v.x, acc[2 * i]);
            acc[2 * i + 1] = fma(c, v.y, acc[2 * i + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < TN; ++i) {
        double nrm = 0.0;
        for (int t = 0; t < T.nnorm; ++t) nrm += T.ncoef[t] * s.P[T.nidx[t]][i];
        ev[i] = acc[i] * T.k / nrm;
    }
}

// the same table evaluated on the isotropic pair (only the b00 n00 term survives), z = qlm_1(1)^2
__device__ __forceinline__ double eval_iso(const TabRef& T, int e, double zr, double zi) {
    return T.iso_coef[e] * (T.iso_im[e] ? zi : zr) * T.k / (T.ncoef[0] * zr);
}

template <int TN>
__global__ void __launch_bounds__(kLanes, 4) eij_orth_kernel(const double2* __restrict__ q1, long long ld1,
                                                             const double2* __restrict__ q2, long long ld2,
                                                             const double2* __restrict__ q3, long long ld3, long long N,
                                                             const double* __restrict__ e1, const double* __restrict__ e2,
                                                             const double* __restrict__ e3, long long lde, OrthCoef K,
                                                             double* __restrict__ Eij, long long ldo) {
    static_assert(TN % 2 == 0, "node tile must be even");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<TN>& s = *reinterpret_cast<Smem<TN>*>(smem_raw);
    const int lane = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * TN;

    // ---- stage the three states (rows 0..14) and the frames
    for (int w = lane; w < 3 * TN * 15; w += kLanes) {
        const int a = w / (TN * 15), r = (w / TN) % 15, i = w % TN;
        const long long p = min(p0 + i, N - 1);             // tail nodes replicate the last one (never stored)
        double2 v = make_double2(0.0, 0.0);
        if (a == 0) v = q1[(long long)r * ld1 + p];
        else if (a == 1) v = q2[(long long)r * ld2 + p];
        else if (q3) v = q3[(long long)r * ld3 + p];
        s.q[a][i][r] = v;
    }
    if (lane < TN * 6) {
        const int i = lane / 6, c = lane % 6;
        const long long p = min(p0 + i, N - 1);
        double E[3][3];
#pragma unroll
        for (int x = 0; x < 3; ++x) { E[0][x] = e1[(long long)x * lde + p]; E[1][x] = e2[(long long)x * lde + p]; E[2][x] = e3[(long long)x * lde + p]; }
        // (v,w): (e1,e1) (e2,e2) (e3,e3) (e2,e3) (e1,e3) (e1,e2)     src/enhancementfactors.f90:147-155
        const int vi = c < 3 ? c : (c == 3 ? 1 : 0), wi = c < 3 ? c : (c == 5 ? 1 : 2);
        double* F = s.F[i][c];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                F[3 * a + b] = E[vi][a] * E[wi][b];                                             // vw = outerprod(v,w)
                F[9 + 3 * a + b] = c < 3 ? ((a == b ? 1.0 / 3.0 : 0.0) - E[vi][a] * E[vi][b])   // tau_vv  :398-405
                                         : (E[vi][a] * E[wi][b] + E[wi][a] * E[vi][b]);         // tau_vw  :407-413
            }
    }
    __syncthreads();
    if (lane < TN * 4) {
        const int i = lane >> 2, a = lane & 3;
        const double thr = (double)1e-8f;
        if (a == 2) s.has3[i] = s.q[2][i][0].x > thr;
        if (a == 3) s.iso3[i] = s.q[0][i][0].x > thr;
        double2 n2[5], n4[9];
        const double2* src = s.q[a == 3 ? 0 : a][i];
#pragma unroll
        for (int m = 0; m < 5; ++m) n2[m] = a == 3 ? make_double2(0.0, 0.0) : src[1 + m];
#pragma unroll
        for (int m = 0; m < 9; ++m) n4[m] = a == 3 ? make_double2(0.0, 0.0) : src[6 + m];
        double u[15];
        sfb::ev_c4_unique(src[0], n2, n4, u);
#pragma unroll
        for (int k = 0; k < 15; ++k) s.U[i][a][k] = u[k];
    }
    make_P(s, 0, 1);
    __syncthreads();

    int any3 = 0, anyno3 = 0, anyiso_no3 = 0;
#pragma unroll
    for (int i = 0; i < TN; ++i) { any3 |= s.has3[i]; anyno3 |= !s.has3[i]; anyiso_no3 |= !s.iso3[i]; }

    // the three tables share one lane <-> entry map, so the lane accumulates, for its entry e and each node, the
    // a4_ii combination (A), the weight of a4_sym2 (S2) and of a4_sym4 (S4); D* are the same for the isotropic state
    const TabRef Tb = tab_c2b2(), Tv = tab_v4(), Tc = tab_c2v2();
    const int e = Tb.perm[lane];
    const bool live = e < 81;
    const int x0 = e % 3, x1 = (e / 3) % 3, x2 = (e / 9) % 3, x3 = e / 27;
    // src/include/ev_c4__body.f90:78: ev(3,2,1,2) aliases ev(1,2,3,3)
    const int uq = !live ? 0 : ((x0 == 2 && x1 == 1 && x2 == 0 && x3 == 1) ? 8 : sfb::a4_unique_index(x0, x1, x2, x3));
    double A[TN], S2[TN], S4[TN], DA[TN], D2[TN], D4[TN], ev[TN];

    // ---- c2b2(q1,q2) = a4_jk(3); a4_ii(1), a4_ii(2) (+ a4_ii(3) when qlm_3 is given)
    eval_table<TN>(Tb, s, lane, ev);
#pragma unroll
    for (int i = 0; i < TN; ++i) {
        A[i] = K.cA[0] * s.U[i][0][uq] + K.cA[1] * s.U[i][1][uq] + (s.has3[i] ? K.cA[2] * s.U[i][2][uq] : 0.0);
        S2[i] = K.cJ2[2] * ev[i];
        S4[i] = K.cJ4[2] * ev[i];
        // isotropic stand-in: every tensor of the "given" branch collapses to a4(iso) / c2b2(iso,iso)
        const double2 b0 = s.q[0][i][0];
        const double zr = b0.x * b0.x - b0.y * b0.y, zi = 2.0 * b0.x * b0.y;
        const double ui = s.U[i][3][uq];
        const double jb = live ? eval_iso(Tb, e, zr, zi) : 0.0;
        if (s.iso3[i]) {
            DA[i] = (K.cA[0] + K.cA[1] + K.cA[2]) * ui;
            D2[i] = (K.cJ2[0] + K.cJ2[1] + K.cJ2[2]) * jb;
            D4[i] = (K.cJ4[0] + K.cJ4[1] + K.cJ4[2]) * jb;
        } else {   // degenerate qlm_1(1) <= 1e-8: the derived-axis closures of the isotropic pair
            const double jc = live ? eval_iso(Tc, e, zr, zi) : 0.0;
            DA[i] = (K.cA[0] + K.cA[1]) * ui + (live ? K.cA[2] * eval_iso(Tv, e, zr, zi) : 0.0);
            D2[i] = K.cJ2[2] * jb + (K.cJ2[0] + K.cJ2[1]) * jc;
            D4[i] = K.cJ4[2] * jb + (K.cJ4[0] + K.cJ4[1]) * jc;
        }
    }
    if (any3) {          // a4_jk(1) = c2b2(q2,q3), a4_jk(2) = c2b2(q1,q3)
        for (int pass = 0; pass < 2; ++pass) {
            __syncthreads();
            make_P(s, pass == 0 ? 1 : 0, 2);
            __syncthreads();
            eval_table<TN>(Tb, s, lane, ev);
#pragma unroll
            for (int i = 0; i < TN; ++i)
                if (s.has3[i]) { S2[i] += K.cJ2[pass] * ev[i]; S4[i] += K.cJ4[pass] * ev[i]; }
        }
    }
    // ---- derived third axis: a4_ii(3) = a4_orth(q1,q2), a4_jk(2) = jointcross(q1,q2), a4_jk(1) = jointcross(q2,q1)
    if (anyno3) {
        if (any3) {
            __syncthreads();
            make_P(s, 0, 1);
            __syncthreads();
        }
        eval_table<TN>(Tv, s, lane, ev);
#pragma unroll
        for (int i = 0; i < TN; ++i)
            if (!s.has3[i]) A[i] += K.cA[2] * ev[i];
        eval_table<TN>(Tc, s, lane, ev);
#pragma unroll
        for (int i = 0; i < TN; ++i)
            if (!s.has3[i]) { S2[i] += K.cJ2[1] * ev[i]; S4[i] += K.cJ4[1] * ev[i]; }
        __syncthreads();
        make_P(s, 1, 0);
        __syncthreads();
        eval_table<TN>(Tc, s, lane, ev);
#pragma unroll
        for (int i = 0; i < TN; ++i)
            if (!s.has3[i]) { S2[i] += K.cJ2[0] * ev[i]; S4[i] += K.cJ4[0] * ev[i]; }
    }
    (void)anyiso_no3;

    // ---- E = Q_lkij vw_kl tau_ji  (doubleinner42 then doubleinner22; src/tensorproducts.f90:153-179):
    // weight the lane's entry for each node and component, reduce over the 81 entries
    const int warp = lane >> 5;
#pragma unroll 1
    for (int ic = 0; ic < TN * 6; ++ic) {
        const int i = ic / 6;
        const double* vw = s.F[0][0] + ic * 18;
        const double* tau = vw + 9;
        auto W = [&](int l, int k, int ii, int jj) { return vw[3 * k + l] * tau[3 * jj + ii]; };
        double num = 0.0, den = 0.0;
        if (live) {
            const double wA = W(x0, x1, x2, x3);
            const double w2 = 0.5 * (wA + W(x2, x3, x0, x1));                                                            // a4_sym2 :53-64
            const double w4 = 0.25 * (W(x0, x2, x1, x3) + W(x2, x1, x0, x3) + W(x0, x3, x2, x1) + W(x3, x1, x2, x0));    // a4_sym4 :66-82
            double a = 0, b2 = 0, b4 = 0, da = 0, d2 = 0, d4 = 0;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (j == i) { a = A[j]; b2 = S2[j]; b4 = S4[j]; da = DA[j]; d2 = D2[j]; d4 = D4[j]; }
            num = wA * a + w2 * b2 + w4 * b4;
            den = wA * da + w2 * d2 + w4 * d4;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            num += __shfl_xor_sync(0xffffffffu, num, o);
            den += __shfl_xor_sync(0xffffffffu, den, o);
        }
        if ((lane & 31) == 0) { s.red[warp][2 * ic] = num; s.red[warp][2 * ic + 1] = den; }
    }
    __syncthreads();
    if (lane < TN * 6) {
        const int i = lane / 6, c = lane % 6;
        const long long p = p0 + i;
        if (p < N) {
            const double num = s.red[0][2 * lane] + s.red[1][2 * lane] + s.red[2][2 * lane];
            const double den = s.red[0][2 * lane + 1] + s.red[1][2 * lane + 1] + s.red[2][2 * lane + 1];
            // n_grain /= 1: the reference's forward rheology silently returns 0, so Evw = 0/0   (homogenizations.f90:316-318)
            Eij[(long long)c * ldo + p] = K.linear ? num / den : nan("");
        }
    }
}

}  // namespace

// Eij_grain = (Ebb, Enn, Evv, Env, Ebv, Enb); alpha is accepted and ignored like the reference (Sachs only).
cudaError_t sfb_launch_eij_orth(const double2* q1, long long ld1, const double2* q2, long long ld2, const double2* q3, long long ld3,
                                long long N, const double* e1, const double* e2, const double* e3, long long lde,
                                const double* Eij_grain, int n_grain, double* Eij, long long ldo, cudaStream_t st) {
    OrthCoef K;
    // src/rheologies.f90:186-206 with n = DFLOAT(n_grain)
    const double n = (double)n_grain;
    double B[6], lam[6], ci[6];
    for (int i = 0; i < 6; ++i) B[i] = pow(Eij_grain[i], 2 / (n + 1));
    lam[0] = -B[0] + B[1] + B[2]; lam[1] = +B[0] - B[1] + B[2]; lam[2] = +B[0] + B[1] - B[2];
    lam[3] = B[3]; lam[4] = B[4]; lam[5] = B[5];
    for (int i = 0; i < 3; ++i) { ci[i] = 4.0 / 3 * lam[i]; ci[3 + i] = 2 * lam[3 + i]; }
    // M2(i) = (a4_ii(ji(i)) + a4_ii(ki(i)) - 2 sym2(a4_jk(i)))/4, ji = [2,3,1], ki = [3,1,2]
    K.cA[0] = (ci[1] + ci[2]) / 4; K.cA[1] = (ci[0] + ci[2]) / 4; K.cA[2] = (ci[0] + ci[1]) / 4;
    for (int i = 0; i < 3; ++i) { K.cJ2[i] = -ci[i] / 2; K.cJ4[i] = ci[3 + i]; }
    K.linear = n_grain == 1;
    // node tile per CTA: 4 by default; SFB_ORTH_TN=8 selects the larger tile (tuning knob, profiles/r01_notes.md)
    static int tn = 0;
    if (!tn) {
        const char* ev = getenv("SFB_ORTH_TN");
        tn = (ev && atoi(ev) == 8) ? 8 : 4;
        cudaError_t e = cudaFuncSetAttribute(eij_orth_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<4>));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(eij_orth_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<8>));
        if (e != cudaSuccess) { tn = 0; return e; }
    }
    const unsigned blocks = (unsigned)((N + tn - 1) / tn);
    if (tn == 8) eij_orth_kernel<8><<<blocks, kLanes, sizeof(Smem<8>), st>>>(q1, ld1, q2, ld2, q3, ld3, N, e1, e2, e3, lde, K, Eij, ldo);
    else eij_orth_kernel<4><<<blocks, kLanes, sizeof(Smem<4>), st>>>(q1, ld1, q2, ld2, q3, ld3, N, e1, e2, e3, lde, K, Eij, ldo);
    return cudaGetLastError();
}
