// Shared declarations of the specfab_b200 CUDA library (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SFB_MAXL 20
#define SFB_NLREG (SFB_MAXL / 2 + 1)

// real(4) constants of the reference promoted to double (oracle/specfab_oracle.py)
#define SFB_PI 3.141592653589793                 // src/header.f90:12
#define SFB_TWOTHIRDS_F 0x1.555556p-1            // 2./3      src/dynamics.f90:576
#define SFB_SQRT2_F 0x1.6a09e6p+0                // sqrt(2.)  src/dynamics.f90:592

struct SfbStepParams {
    const double2* nlm_in;   // [n][ld_in]  complex(8), node contiguous (Fortran nlm(N,n))
    double2* nlm_out;        // [n][ld_out]
    const double* ugrad;     // [9][ld_u]   Fortran ugrad(N,3,3): plane p = i + 3 j
    const double* tau;       // [9][ld_t]   stress (DDRX); may be null -> tau := D (sym part of ugrad)
    const double* gamma0_arr;  // optional per-node DDRX rate factor [N] (null -> gamma0)
    const double* lambda_arr;  // optional per-node CDRX rate factor [N] (null -> lambda)
    long long N, ld_in, ld_out, ld_u, ld_t;
    double dt, iota, zeta, nu_mult, gamma0, lambda;
    int nstage;              // 1 = Euler, 4 = classical RK4
    int use_lrot, use_reg;
    const double2* ktab;     // set by the launcher in gtab mode
    int raw_ok;              // set by the launcher: forcing planes can be staged with 1-D TMA bulk copies
    int n0_global;           // set by the launcher: RK4 re-reads n0 from global (3 smem buffers)
    int rio;                 // reduced I/O (sfb_step_rnlm_arr): nlm_in / nlm_out hold the rows m >= 0 only, row (l, m) at
                             // (l/2)^2 + m (src/reducedform.f90:160-187); reduced kernels only
};

// per-L regularisation constants, set by sfb_init (host pow(), like the reference's libm)
struct SfbRegConst {
    double nu;                 // src/include/regcalib.f90
    double regdiag[SFB_NLREG]; // abs(l(l+1)/(L(L+1)))**expo, l = 0,2,..,L   src/dynamics.f90:512
};

typedef cudaError_t (*sfb_step_launch_fn)(const SfbStepParams&, const SfbRegConst&, cudaStream_t);

// Per (device, stream) queue of tiles a fast step kernel hands to its general-state twin (sfb_step_wloop.cuh); owned by
// sfb_api.cu, grown on demand (synchronises the stream when it has to reallocate), released by sfb_finalize.
// One int flag per tile, rewritten by the fast kernel on every launch.
cudaError_t sfb_worklist_get(cudaStream_t st, long long ntile, int** buf);
void sfb_worklist_release();
