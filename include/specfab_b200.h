/* specfab_b200 -- C ABI of the B200-native batched fabric-evolution engine.
 *
 * Drop-in boundary for the hot path of nicholasmr/specfab.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference tree).
 * Plain C: pointers + sizes, no C++/torch types.  Error handling: every function returns
 * SFB_OK (0) or a negative SFB_E* code; sfb_last_error() gives the message.  (The reference
 * `stop`s the process instead: src/homogenizations.f90:95,112,183,220; src/frames.f90:48.)
 *
 * Array convention (same as f2py hands to Fortran): batched arguments are the
 * reference's own arrays with a LEADING node dimension in Fortran column-major order, i.e. the
 * node index is the contiguous one:
 *     nlm(N, nlm_len) complex(8)  ->  double[2*ld*nlm_len], element (node p, coef j) at 2*(j*ld+p)
 *     ugrad(N,3,3), tau(N,3,3)    ->  double[9*ld],         element (p,i,k) at (i+3k)*ld + p
 *     Eij(N,6), lami(N,3), ei(N,3,3) [ei(p,i,:) = i-th eigenvector], a2(N,3,3), a4(N,3,3,3,3)
 * `ld` (leading dimension, >= N) is given per call.
 *
 * Pointer spaces: functions ending in `_dev` take DEVICE pointers and a cudaStream_t (passed as
 * void*, may be NULL) and are asynchronous; the others take HOST pointers, stage through device
 * memory on the current CUDA device and return when the result is in the output buffers.
 * There is no CPU fallback: without a usable CUDA device every call returns SFB_ECUDA.
 *
 * Threading (the reference is single threaded and not re-entrant: module globals src/header.f90:16, SAVE'd locals
 * src/homogenizations.f90:80).  sfb_init / sfb_finalize / sfb_set_variant change process-wide state and must not run
 * concurrently with any other call.  Between them every entry point may be called from several host threads on distinct
 * buffers: `_dev` calls are re-entrant (give each thread its own stream); host-pointer calls on the same device take turns
 * on that device's staging ring, host-pointer calls on different devices (sfb_set_device per thread, or
 * sfb_step_arr_multi) run concurrently.  sfb_last_error() reports the calling thread's last failure.
 */
#ifndef SPECFAB_B200_H
#define SPECFAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SFB_OK = 0,
    SFB_EINVAL = -1,   /* bad argument (odd L, L<4, L>20, null pointer, ld<N ...) */
    SFB_ENOINIT = -2,  /* sfb_init() not called */
    SFB_ECUDA = -3,    /* CUDA runtime error / no device */
    SFB_ENOTBUILT = -4 /* kernel for this L was not compiled into the library */
};

/* term mask of the fused step (src/specfabpy/integrator.py:55-77) */
enum { SFB_LROT = 1, SFB_DDRX = 2, SFB_CDRX = 4, SFB_REG = 8 };
/* time integrators: Euler = src/dynamics.f90:108 ; RK4 = classical, BASELINE config 2 */
enum { SFB_EULER = 1, SFB_RK4 = 4 };

/* per-node status flags written by the Eij routines (0 = ok) */
enum { SFB_ST_TAYLOR_FALLBACK = 1, SFB_ST_TAYLOR_FAILED = 2, SFB_ST_NONFINITE = 4 };

/* init(L) -> lm(2,nlm_len), nlm_len        src/specfabpy.f90:150-160, src/specfab.f90:37-53
 * Loads tables, fixes L for subsequent calls (module state in the reference).  Idempotent. */
int sfb_init(int L);
int sfb_nlm_len(void);                 /* src/specfabpy.f90:162-166 */
int sfb_get_lm(int32_t* lm /* [2*nlm_len], lm(1,j)=l, lm(2,j)=m */);
void sfb_finalize(void);
const char* sfb_last_error(void);
/* library/build information: JSON string (compiled L list, tile shapes, DFMA counts) */
const char* sfb_build_info(void);

typedef struct sfb_step_opts {
    double dt;
    double iota;        /* M_LROT eps^1 coefficient       src/dynamics.f90:52 */
    double zeta;        /* M_LROT eps^2 coefficient */
    double nu_mult;     /* multiplier on M_REG (1 = calibrated regularisation) */
    double gamma0;      /* DDRX rate factor (caller-multiplied in the reference, src/dynamics.f90:260) */
    double lambda;      /* CDRX rate factor (src/dynamics.f90:483) */
    const double* gamma0_arr; /* optional per-node rate factors [N] (same pointer space as the call), or NULL */
    const double* lambda_arr;
    int32_t terms;      /* SFB_LROT | SFB_DDRX | SFB_CDRX | SFB_REG */
    int32_t scheme;     /* SFB_EULER | SFB_RK4 */
    int32_t nsteps;     /* number of consecutive steps with the same forcing (>= 1) */
    int32_t reserved;   /* flags: 0, or SFB_STEP_GENERAL (see sfb_step_arr) */
} sfb_step_opts;

/* Fused batched time step:  nlm <- nlm + dt * (M_LROT + gamma0*M_DDRX + lambda*M_CDRX + M_REG) nlm
 * Replaces the per-node loop  M = sf.M_LROT(..) + sf.M_REG(..) + ..; nlm + dt*matmul(M,nlm)
 * (src/specfabpy/integrator.py:73-77, src/dynamics.f90:99-110, src/specfabpy/fenics/CPO.py:200-202).
 * D, W are the symmetric / antisymmetric parts of ugrad; tau may be NULL (then tau := D, as
 * src/specfabpy/integrator.py:39).  nlm_in may equal nlm_out (in-place).
 *
 * Real-valued ODFs and general complex vectors.  The reference operators are plain linear maps on arbitrary complex vectors.
 * Every physical state has the real-ODF symmetry n_l^-m = (-1)^m conj(n_l^m), Im n_l^0 = 0, which the operators preserve, and
 * the default kernels exploit it: a 32-node tile whose input has the symmetry to round-off (every component of the defect
 * within 2^-46 |n_0^0|) is advanced from its rows m >= 0 only and its output is EXACTLY symmetric (the defect, <= 1.4e-14
 * relative, is projected out); a tile with any node outside that bound takes the general path, which keeps an antisymmetric
 * part.  A node with a defect below the bound can therefore differ at the 1e-14 level depending on which nodes share its tile
 * (batch order, chunking, device split).  Callers that need results independent of the batching -- or that deliberately
 * carry a tiny antisymmetric part -- set SFB_STEP_GENERAL in sfb_step_opts.reserved: every tile then takes the general path. */
#define SFB_STEP_GENERAL 1   /* sfb_step_opts.reserved: treat every state as a general complex vector */
int sfb_step_arr(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld,
                 const double* ugrad, const double* tau, const sfb_step_opts* opts);
int sfb_step_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                     const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                     const sfb_step_opts* opts, void* stream);

/* One host array on SEVERAL GPUs of this process (SURVEY.md 8b "multi-GPU selection via explicit device list", 8e): the nodes
 * are split into contiguous ranges on 32-node boundaries, range i goes to devices[i]; one host thread, one staging ring and
 * three streams per device, no inter-GPU traffic (every node is independent: src/dynamics.f90:52,251, the per-node loops of
 * src/specfabpy.f90:483-485).  This is the entry point for the reference's single-process callers (Fortran programs, f2py,
 * an Elmer rank): same arguments as sfb_step_arr plus the device ordinals.  The result is bit-identical to the one-device
 * call; the calling thread's current device is restored.  sfb_init must have been called once (it holds no per-device state
 * the step needs). */
int sfb_step_arr_multi(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld, const double* ugrad, const double* tau,
                       const sfb_step_opts* opts, const int* devices, int ndev);
int sfb_step_rnlm_arr_multi(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld, const double* ugrad, const double* tau,
                            const sfb_step_opts* opts, const int* devices, int ndev);

/* The same fused step on REDUCED-FORM states: rnlm(N, rnlm_len) holds the m >= 0 coefficients of a real-valued ODF, row
 * (l, m) at (l/2)^2 + m -- the layout of nlm_to_rnlm / rnlm_to_nlm (src/reducedform.f90:160-187, src/specfabpy.f90:1076-1094),
 * the representation the FE couplers keep their state in (src/specfabpy/fenics/CPO.py:103-118, 339-365).  Equal to
 * nlm_to_rnlm(step(rnlm_to_nlm(rnlm))) bit for bit, but only (L+2)^2/4 instead of (L+1)(L+2)/2 coefficient rows are read,
 * written and (host variant) cross PCIe; Im n_l^0 is taken as 0.  ld = leading (node) dimension of the rnlm arrays. */
int sfb_step_rnlm_arr(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld,
                      const double* ugrad, const double* tau, const sfb_step_opts* opts);
int sfb_step_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                          const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t,
                          const sfb_step_opts* opts, void* stream);

/* One call per FE time step (SURVEY.md 8b "step_moments_Eij_arr", BASELINE config 5): the fused step above, then on the
 * new state a2 (optional), a4 (optional), the a2 eigenframe (ei, lami: optional, both or none) and the eigenenhancements
 * Eij (N,6) in that frame -- what src/specfabpy/fenics/CPO.py:evolve + src/specfabpy/fenics/enhancementfactor.py:101-128 do
 * per node.  All launches go to `stream` back to back; the state never leaves the device between them.  Any of
 * a2/a4/ei/lami/status may be NULL.  Device pointers; arrays are packed with leading dimension N except nlm (ld). */
int sfb_step_moments_Eij_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                                 const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t, const sfb_step_opts* opts,
                                 const double* Eij_grain, double alpha, int n_grain,
                                 double* Eij, double* a2, double* a4, double* ei, double* lami, int32_t* status, void* stream);

/* The FE time step for couplers that keep the state in reduced form (src/specfabpy/fenics/CPO.py): step_rnlm, then a2
 * (optional), the a2 eigenframe (optional) and Eij of the new state, all read from the m >= 0 rows.  n_grain = 1 or -3.
 * sfb_Eij_eigenframe_rnlm_arr_dev is the second half alone (a2 -> eigenframe -> Eij of a reduced-form state). */
int sfb_Eij_eigenframe_rnlm_arr_dev(const double* rnlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                                    double* Eij, double* a2, double* ei, double* lami, int32_t* status, void* stream);
int sfb_step_moments_Eij_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out,
                                      const double* ugrad, int64_t ld_u, const double* tau, int64_t ld_t, const sfb_step_opts* opts,
                                      const double* Eij_grain, double alpha, int n_grain,
                                      double* Eij, double* a2, double* ei, double* lami, int32_t* status, void* stream);

/* a2(nlm) -> (N,3,3)                         src/specfabpy.f90:583-590, src/moments.f90:37-44 */
int sfb_a2_arr(const double* nlm, int64_t N, int64_t ld, double* a2);
int sfb_a2_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a2, void* stream);
/* a4(nlm) -> (N,3,3,3,3)                     src/specfabpy.f90:592-599, src/moments.f90:46-55.
 * Reproduces the reference's real(4) constants and its ev(3,2,1,2) alias quirk (ev_c4__body.f90:78). */
int sfb_a4_arr(const double* nlm, int64_t N, int64_t ld, double* a4);
int sfb_a4_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a4, void* stream);
/* eig(nlm) -> ei(N,3,3), lami(N,3): eigenframe of a2, largest eigenvalue first
 *                                            src/specfabpy.f90:312-320, src/frames.f90:14-22.
 * Eigenvector sign: largest |component| positive (LAPACK's choice is unpinned). */
int sfb_eig_arr(const double* nlm, int64_t N, int64_t ld, double* ei, double* lami);
int sfb_eig_arr_dev(const double* nlm, int64_t N, int64_t ld, double* ei, double* lami, void* stream);
/* eigframe_arr(M(N,3,3), plane in {"ij","xy","xz"})   src/specfabpy.f90:333-344, src/frames.f90:24-60 */
int sfb_eigframe_arr(const double* M, int64_t N, const char* plane, double* ei, double* lami);
int sfb_eigframe_arr_dev(const double* M, int64_t N, int64_t ld, const char* plane, double* ei, double* lami, void* stream);
/* Eij_tranisotropic_arr(nlm, e1,e2,e3 (N,3), Eij_grain(2), alpha, n_grain) -> Eij(N,6) = (E11,E22,E33,E23,E13,E12)
 *                                            src/specfabpy.f90:474-486, src/enhancementfactors.f90:23-69.
 * n_grain = 1, 3 (Sachs through a6/a8, needs L >= 8; src/homogenizations.f90:93-102) or -3 (:105-109); the Taylor
 * part is the n'=1 solve in every case like the reference (:148).  status (optional, [N]) receives SFB_ST_* flags
 * instead of the reference's `stop` (src/homogenizations.f90:183). */
int sfb_Eij_tranisotropic_arr(const double* nlm, int64_t N, int64_t ld, const double* e1, const double* e2, const double* e3,
                              const double* Eij_grain, double alpha, int n_grain, double* Eij, int32_t* status);
int sfb_Eij_tranisotropic_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* e1, const double* e2, const double* e3,
                                  const double* Eij_grain, double alpha, int n_grain, double* Eij, int32_t* status, void* stream);
/* Evw_tranisotropic(nlm, v, w, tau, Eij_grain, alpha, n_grain) -> Evw, batched over nodes: v, w (N,3), tau (N,3,3), Evw (N)
 *                                            src/specfabpy.f90:379-388, src/enhancementfactors.f90:47-69.
 * The generalized enhancement factor for an ARBITRARY direction pair and stress (Eij_tranisotropic is its six eigenframe
 * pairs with tau_vv / tau_vw, src/enhancementfactors.f90:398-413).  n_grain = 1 or -3 (n' = 3 through Eij_tranisotropic_arr). */
int sfb_Evw_tranisotropic_arr(const double* nlm, int64_t N, int64_t ld, const double* v, const double* w, const double* tau,
                              const double* Eij_grain, double alpha, int n_grain, double* Evw, int32_t* status);
int sfb_Evw_tranisotropic_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* v, const double* w, const double* tau,
                                  const double* Eij_grain, double alpha, int n_grain, double* Evw, int32_t* status, void* stream);
/* Fused a2 -> eigenframe -> Eij_tranisotropic in that frame (the eigenenhancements); batches
 * src/specfabpy/fenics/enhancementfactor.py:101-128 / src/specfabpy/common.py:13-37.  ei/lami optional outputs. */
int sfb_Eij_eigenframe_arr(const double* nlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                           double* Eij, double* ei, double* lami, int32_t* status);
int sfb_Eij_eigenframe_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* Eij_grain, double alpha, int n_grain,
                               double* Eij, double* ei, double* lami, int32_t* status, void* stream);

/* a6 of every node -> (N,3,3,3,3,3,3) Fortran order (needs L >= 6)          src/specfabpy.f90:601-608, src/moments.f90:57-66 */
int sfb_a6_arr(const double* nlm, int64_t N, int64_t ld, double* a6);
int sfb_a6_arr_dev(const double* nlm, int64_t N, int64_t ld, double* a6, void* stream);
/* E_CAFFE_arr(nlm, eps (N,3,3), Emin, Emax, n_grain) -> E(N)                 src/specfabpy.f90:543-554,
 * src/enhancementfactors.f90:301-331.  n_grain = 3 uses <D> from RSS^4 (ev_D4, needs L >= 8), anything else RSS^2. */
int sfb_E_CAFFE_arr(const double* nlm, int64_t N, int64_t ld, const double* eps, double Emin, double Emax, int n_grain, double* E);
int sfb_E_CAFFE_arr_dev(const double* nlm, int64_t N, int64_t ld, const double* eps, double Emin, double Emax, int n_grain, double* E,
                        void* stream);
/* pfJ(nlm, Lmax) of every node -> J(N): pole-figure J index truncated at Lmax <= L     src/specfabpy.f90:729-736,
 * src/idealstate.f90:111-124 */
int sfb_pfJ_arr(const double* nlm, int64_t N, int64_t ld, int Lmax, double* J);
int sfb_pfJ_arr_dev(const double* nlm, int64_t N, int64_t ld, int Lmax, double* J, void* stream);
/* State ingest: a2 (N,3,3) -> nlm (N,6);  a4 (N,3,3,3,3) -> nlm (N,15);  a6 (N,3^6) -> nlm (N,28)   (rank = 2, 4, 6)
 *                                            src/specfabpy.f90:619-647, src/moments.f90:68-92.
 * The output holds the l <= rank coefficients; the caller embeds them in a longer state (higher l = 0). */
int sfb_ai_to_nlm_arr(int rank, const double* a, int64_t N, double* nlm);
int sfb_ai_to_nlm_arr_dev(int rank, const double* a, int64_t N, int64_t ld, double* nlm, int64_t ld_nlm, void* stream);
/* Eij_orthotropic_arr(nlm_1, nlm_2, nlm_3 (N,nlm_len), e1,e2,e3 (N,3), Eij_grain(6), alpha, n_grain) -> Eij(N,6)
 *                                            src/specfabpy.f90:488-500, src/enhancementfactors.f90:134-189.
 * Orthotropic grains (olivine): nlm_1..3 are the distributions of the slip-system axes (b, n, v); where
 * REAL(nlm_3(1)) <= 1e-8 (or nlm_3 == NULL) the third axis is derived from the first two
 * (src/moments.f90:357-384).  Eij_grain = (Ebb, Enn, Evv, Env, Ebv, Enb).  Like the reference only the Sachs
 * bound is evaluated (alpha is ignored) and only n_grain = 1 is supported: any other n_grain yields NaN, the
 * reference's 0/0 (src/homogenizations.f90:316-318).  Each state has its own leading dimension. */
int sfb_Eij_orthotropic_arr(const double* nlm_1, const double* nlm_2, const double* nlm_3, int64_t N, int64_t ld,
                            const double* e1, const double* e2, const double* e3,
                            const double* Eij_grain, double alpha, int n_grain, double* Eij);
int sfb_Eij_orthotropic_arr_dev(const double* nlm_1, int64_t ld1, const double* nlm_2, int64_t ld2, const double* nlm_3, int64_t ld3,
                                int64_t N, const double* e1, const double* e2, const double* e3,
                                const double* Eij_grain, double alpha, int n_grain, double* Eij, void* stream);

/* Batched operator export (the form the Eulerian FE couplers consume: src/specfabpy/fenics/CPO.py:200-202,
 * src/specfabpy/firedrake/ice.py:202-204).  M is (N, nlm_len, nlm_len) in Fortran order, complex(8) for
 * M_LROT / M_DDRX(_src), real(8) for M_REG.  eps, omg, tau: (N,3,3).
 *   M_LROT(nlm, eps, omg, iota, zeta)      src/specfabpy.f90:172-180, src/dynamics.f90:52-97
 *   M_DDRX_src(nlm, tau)                   src/specfabpy.f90:192-199, src/dynamics.f90:277-298
 *   M_DDRX(nlm, tau) = M_DDRX_src - <D> I  src/specfabpy.f90:182-190, src/dynamics.f90:251-275
 *   M_REG(nlm, eps)                        src/specfabpy.f90:237-245, src/dynamics.f90:494-518
 *   M_CDRX(nlm) = diag(-l(l+1))            src/specfabpy.f90:228-235, src/dynamics.f90:474-492 (constant: host only) */
int sfb_M_LROT_arr(const double* eps, const double* omg, int64_t N, double iota, double zeta, double* M);
int sfb_M_LROT_arr_dev(const double* eps, const double* omg, int64_t N, int64_t ld, double iota, double zeta, double* M, void* stream);
int sfb_M_DDRX_src_arr(const double* tau, int64_t N, double* M);
int sfb_M_DDRX_src_arr_dev(const double* tau, int64_t N, int64_t ld, double* M, void* stream);
int sfb_M_DDRX_arr(const double* nlm, int64_t ld_nlm, const double* tau, int64_t N, double* M);
int sfb_M_DDRX_arr_dev(const double* nlm, int64_t ld_nlm, const double* tau, int64_t N, int64_t ld, double* M, void* stream);
int sfb_M_REG_arr(const double* eps, int64_t N, double* M);
int sfb_M_REG_arr_dev(const double* eps, int64_t N, int64_t ld, double* M, void* stream);
int sfb_M_CDRX(double* M /* [nlm_len*nlm_len] */);

/* Reduced-form operators (src/reducedform.f90:76-120, specfabpy reduce_M src/specfabpy.f90:1066-1074): with rnlm the m >= 0
 * coefficients,  d(rnlm)/dt = Mrr Re(rnlm) + Mri Im(rnlm) + i (Mir Re(rnlm) + Mii Im(rnlm)).  Every output is
 * (N, rnlm_len, rnlm_len) real(8), Fortran order.  This is the form the FE couplers assemble
 * (src/specfabpy/fenics/CPO.py:200-211, src/specfabpy/firedrake/ice.py:202-210).
 *   sfb_reduce_M_arr           reduce an existing dense operator M (N, nlm_len, nlm_len), complex(8) or real(8)
 *   sfb_M_LROT_reduced_arr     reduce_M(M_LROT(nlm, eps, omg, iota, zeta)) without materialising the dense M
 *   sfb_M_DDRX_reduced_arr     reduce_M(M_DDRX(nlm, tau)), or of M_DDRX_src(tau) when src_only != 0 (nlm may then be NULL) */
int sfb_reduce_M_arr(const double* M, int is_complex, int64_t N, double* Mrr, double* Mri, double* Mir, double* Mii);
int sfb_reduce_M_arr_dev(const double* M, int is_complex, int64_t N, int64_t ld, double* Mrr, double* Mri, double* Mir, double* Mii,
                         void* stream);
int sfb_M_LROT_reduced_arr(const double* eps, const double* omg, int64_t N, double iota, double zeta,
                           double* Mrr, double* Mri, double* Mir, double* Mii);
int sfb_M_LROT_reduced_arr_dev(const double* eps, const double* omg, int64_t N, int64_t ld, double iota, double zeta,
                               double* Mrr, double* Mri, double* Mir, double* Mii, void* stream);
int sfb_M_DDRX_reduced_arr(const double* nlm, int64_t ld_nlm, const double* tau, int64_t N, int src_only,
                           double* Mrr, double* Mri, double* Mir, double* Mii);
int sfb_M_DDRX_reduced_arr_dev(const double* nlm, int64_t ld_nlm, const double* tau, int64_t N, int64_t ld, int src_only,
                               double* Mrr, double* Mri, double* Mir, double* Mii, void* stream);

/* apply_bounds(nlm): rescale the l=2 / l=4 blocks whose power spectrum exceeds the delta-function bound
 *                                            src/specfabpy.f90:764-771, src/dynamics.f90:530-557 */
int sfb_apply_bounds_arr(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld);
int sfb_apply_bounds_arr_dev(const double* nlm_in, double* nlm_out, int64_t N, int64_t ld_in, int64_t ld_out, void* stream);
/* reduced form of real-valued ODFs: rnlm(N, rnlm_len) holds the m >= 0 coefficients, (L+2)^2/4 of them
 *                                            src/specfabpy.f90:1039-1061, src/reducedform.f90:160-187 */
/* apply_bounds on a reduced-form field, device pointers: what src/specfabpy/fenics/CPO.py:339-365 does per node through
 * rnlm_to_nlm / apply_bounds / nlm_to_rnlm (same rescaling factors as sfb_apply_bounds_arr, bit for bit). */
int sfb_apply_bounds_rnlm_arr_dev(const double* rnlm_in, double* rnlm_out, int64_t N, int64_t ld_in, int64_t ld_out, void* stream);
int sfb_rnlm_len(void);
int sfb_nlm_to_rnlm_arr(const double* nlm, double* rnlm, int64_t N);
int sfb_nlm_to_rnlm_arr_dev(const double* nlm, double* rnlm, int64_t N, int64_t ld_nlm, int64_t ld_rnlm, void* stream);
int sfb_rnlm_to_nlm_arr(const double* rnlm, double* nlm, int64_t N);
int sfb_rnlm_to_nlm_arr_dev(const double* rnlm, double* nlm, int64_t N, int64_t ld_rnlm, int64_t ld_nlm, void* stream);

/* tuning knob: select an alternative compiled kernel variant (0 = default); unknown ids fall back to 0 */
int sfb_set_variant(int variant);

/* simple device-memory helpers so that FFI callers need no CUDA binding of their own */
int sfb_dev_malloc(void** p, int64_t bytes);
int sfb_dev_free(void* p);
int sfb_memcpy_h2d(void* dst, const void* src, int64_t bytes);
int sfb_memcpy_d2h(void* dst, const void* src, int64_t bytes);
int sfb_host_alloc_pinned(void** p, int64_t bytes);
int sfb_host_free_pinned(void* p);
/* Page-lock an EXISTING host array (a Fortran allocatable, a numpy array) so that the host-pointer entry points move it by
 * DMA at PCIe speed: sfb_step_arr on pageable memory is staged by the driver and ~4x slower (1.3e7 vs 5.6e7 node-updates/s,
 * BASELINE config 2 on B200).  Register once, reuse every step, unregister before the array is freed. */
int sfb_host_register(void* p, int64_t bytes);
int sfb_host_unregister(void* p);
int sfb_sync(void);
int sfb_device_count(void);
int sfb_set_device(int dev);

#ifdef __cplusplus
}
#endif
#endif
