// Device helpers shared by the step-kernel skeletons: quadric coefficients, DDRX weights and the
// per-node forcing preparation.  Included INSIDE the anonymous namespace of a step translation unit,
// after kTN, kNF, c_reg and the SC_* scalar slots are defined.
#include "sfb_step_math.cuh"

// ---------------------------------------------------------------------------------------------
// Per-node forcing preparation, split into independent tasks so that the threads of a CTA share it
// (task = tid / kTN is warp uniform for kTN >= 32):
//   task 0: M_LROT weights + M_REG / M_CDRX scalars      task 1: DDRX weights g      task 2: <D> ingredients
// ---------------------------------------------------------------------------------------------
// where a node's 3x3 forcing matrices are read from: global arrays (stride = leading dimension) or the
// TMA-staged shared-memory copy of the tile (stride = kTN); element (i,k) at base[(i+3k)*stride]
struct ForcSrc {
    const double* ug; long long su;
    const double* tau; long long st;      // tau == nullptr -> tau := D
};
__device__ __forceinline__ ForcSrc global_src(const SfbStepParams& P, long long node) {
    ForcSrc s; s.ug = P.ugrad + node; s.su = P.ld_u; s.tau = P.tau ? P.tau + node : nullptr; s.st = P.ld_t; return s;
}

__device__ __forceinline__ void load_sym_skew(const ForcSrc& S, double D[3][3], double W[3][3]) {
    double u[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) u[i][j] = S.ug[(long long)(i + 3 * j) * S.su];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { D[i][j] = (u[i][j] + u[j][i]) / 2; W[i][j] = (u[i][j] - u[j][i]) / 2; }
}

__device__ __forceinline__ void load_tau(const ForcSrc& S, double T[3][3]) {
    if (S.tau) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) T[i][j] = S.tau[(long long)(i + 3 * j) * S.st];
    } else {   // tau := D (src/specfabpy/integrator.py:39)
        double W[3][3];
        load_sym_skew(S, T, W);
    }
}

// TNS: node stride of the forcing / scalar arrays; WITHB: also fill the mirrored block of lane set B
// (the one-lane reduced kernel, sfb_step_kernel_r.cuh, needs set A only)
template <int TNS = kTN, bool WITHB = true>
__device__ void prep_lrot(const SfbStepParams& P, const ForcSrc& S, long long node, int t, double2* forc, double* scal,
                          const double* upre = nullptr) {      // upre: the node's ugrad (element (i,k) at i + 3k) when the caller loaded it early
    constexpr int kTN = TNS;                   // shadows the tile constant inside this function
    double D[3][3], W[3][3];
    if (upre) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { D[i][j] = (upre[i + 3 * j] + upre[j + 3 * i]) / 2; W[i][j] = (upre[i + 3 * j] - upre[j + 3 * i]) / 2; }
    } else {
        load_sym_skew(S, D, W);
    }
    double2* fA = forc + t;                    // lane set A block: entry e at fA[e*kTN]
    double2* fB = forc + kNF * kTN + t;        // lane set B block
    // ---- M_LROT weights, src/dynamics.f90:71-76
    double sq[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) sq[i][j] = D[i][0] * D[0][j] + D[i][1] * D[1][j] + D[i][2] * D[2][j];
    const double zetanorm = P.zeta / sqrt(sq[0][0] + sq[1][1] + sq[2][2]);   // 0/0 -> NaN like the reference
    double E[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) E[i][j] = P.iota * D[i][j] + zetanorm * sq[i][j];
    double2 qe[5], qo[3];
    quad_rr(E, qe);
    quad_tp(W, qo);
    if (!P.use_lrot) {
#pragma unroll
        for (int i = 0; i < 5; ++i) qe[i] = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 3; ++i) qo[i] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int d = -2; d <= 2; ++d) { fA[(d + 2) * kTN] = qe[d + 2]; if (WITHB) fB[(d + 2) * kTN] = qe[-d + 2]; }
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
        fA[(5 + d + 1) * kTN] = make_double2(-qo[d + 1].y, qo[d + 1].x);                   //  i*qo[d]
        if (WITHB) fB[(5 + d + 1) * kTN] = make_double2(qo[-d + 1].y, -qo[-d + 1].x);      // -i*qo[-d]
    }
    // M_REG: -nu*||D||_F * regdiag   src/dynamics.f90:516-517
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) fro += D[i][j] * D[i][j];
    scal[SC_RM * kTN + t] = P.use_reg ? -(P.nu_mult * (c_reg.nu * sqrt(fro))) : 0.0;
    scal[SC_LAM * kTN + t] = P.lambda_arr ? P.lambda_arr[node] : P.lambda;
    scal[SC_C0 * kTN + t] = 0.0;
}

#if SFB_DDRX
template <int TNS = kTN, bool WITHB = true>
__device__ void prep_ddrx_g(const SfbStepParams& P, const ForcSrc& S, long long node, int t, double2* forc, double* scal) {
    constexpr int kTN = TNS;
    double2* fA = forc + t;
    double2* fB = forc + kNF * kTN + t;
    const double g0 = P.gamma0_arr ? P.gamma0_arr[node] : P.gamma0;
    scal[SC_G0 * kTN + t] = g0;
    double T[3][3];
    load_tau(S, T);
    double2 qt[5], g[15];
    quad_rr(T, qt);
    ddrx_weights_raw(qt, g);
    double dd = 0.0;                                   // doubleinner22(tau,tau) = tau_ij tau_ji
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dd += T[i][j] * T[j][i];
    const double kk = 0x1.14d2dcd9ceb17p-3;            // (3*Sqrt(5/Pi))/28.
    // g = k*g*5/(tau:tau) (src/dynamics.f90:293), then *Gamma0: one division per node instead of 30
    const double sc = g0 * ((kk * 5) / dd);
#pragma unroll
    for (int k = 0; k < 15; ++k) g[k] = make_double2(sc * g[k].x, sc * g[k].y);
#pragma unroll
    for (int k = 0; k < 15; ++k) { fA[(8 + k) * kTN] = g[k]; if (WITHB) fB[(8 + k) * kTN] = g[cat_mirror(k)]; }
}

template <int TNS = kTN>
__device__ void prep_ddrx_d(const ForcSrc& S, int t, double* scal) {
    constexpr int kTN = TNS;
    // <D> ingredients, src/dynamics.f90:415-417
    double T[3][3];
    load_tau(S, T);
    double sq[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) sq[i][j] = T[i][0] * T[0][j] + T[i][1] * T[1][j] + T[i][2] * T[2][j];
    const double tv[6] = {T[0][0], T[1][1], T[2][2], SFB_SQRT2 * T[1][2], SFB_SQRT2 * T[0][2], SFB_SQRT2 * T[0][1]};
    const double sv[6] = {sq[0][0], sq[1][1], sq[2][2], SFB_SQRT2 * sq[1][2], SFB_SQRT2 * sq[0][2], SFB_SQRT2 * sq[0][1]};
#pragma unroll
    for (int p = 0; p < 6; ++p) { scal[(SC_TAUV + p) * kTN + t] = tv[p]; scal[(SC_TSQV + p) * kTN + t] = sv[p]; }
    scal[SC_NORM * kTN + t] = sq[0][0] + sq[1][1] + sq[2][2];
}
#endif

// <D> of one node (src/dynamics.f90:402-422) with its ingredients (tau, tau.tau, tau:tau) recomputed from the forcing
// array: 9 loads and ~40 flops per node and stage buy back 13 doubles of shared memory per node (SFB_NO_DSCAL skeletons)
__device__ __forceinline__ double ddrx_davg(const ForcSrc& S, double2 n00, const double2 n2[3], const double2 n4[5]) {
    double T[3][3];
    load_tau(S, T);
    double sq[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) sq[i][j] = T[i][0] * T[0][j] + T[i][1] * T[1][j] + T[i][2] * T[2][j];
    const double tv[6] = {T[0][0], T[1][1], T[2][2], SFB_SQRT2 * T[1][2], SFB_SQRT2 * T[0][2], SFB_SQRT2 * T[0][1]};
    const double sv[6] = {sq[0][0], sq[1][1], sq[2][2], SFB_SQRT2 * sq[1][2], SFB_SQRT2 * sq[0][2], SFB_SQRT2 * sq[0][1]};
    return sfb::ev_D2(n00, n2, n4, tv, sv, sq[0][0] + sq[1][1] + sq[2][2]);
}

// task dispatch: kThreads >= 2*kTN in both skeletons
__device__ __forceinline__ void prep_tile(const SfbStepParams& P, long long node0, int nvalid, int tid, double2* forc, double* scal) {
    const int task = tid / kTN, t = tid - task * kTN;
    if (t >= nvalid) return;
    const ForcSrc S = global_src(P, node0 + t);
    if (task == 0) prep_lrot(P, S, node0 + t, t, forc, scal);
#if SFB_DDRX
#ifdef SFB_NO_DSCAL
    if (task == 1) prep_ddrx_g(P, S, node0 + t, t, forc, scal);
#else
    if (kThreads >= 3 * kTN) {
        if (task == 1) prep_ddrx_g(P, S, node0 + t, t, forc, scal);
        if (task == 2) prep_ddrx_d(S, t, scal);
    } else if (task == 1) {
        prep_ddrx_g(P, S, node0 + t, t, forc, scal);
        prep_ddrx_d(S, t, scal);
    }
#endif
#endif
}
