/* Plain-C caller of the C ABI (include/specfab_b200.h): what a cgo / JNI / iso_c_binding binding sees.
 * usage: abi_smoke <N>   -- one Euler LROT+REG step of N isotropic nodes under uniaxial compression (host pointers),
 * prints the return code of sfb_init and, when a device is present, the new state as hex doubles.
 * usage: abi_smoke <N> multi -- the same step through sfb_step_arr_multi on every device of the box (what a single-process
 * Fortran / Elmer caller would use); the output must be bit-identical. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "specfab_b200.h"

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 3;
    int rc = sfb_init(8);
    printf("init %d\n", rc);
    if (rc != SFB_OK) {
        printf("error %s\n", sfb_last_error());
        return rc == SFB_ECUDA ? 0 : 1;     /* no device: the library must say so, not fall back */
    }
    const int n = sfb_nlm_len();
    double* x = (double*)calloc((size_t)(2 * n * N), sizeof(double));
    double* y = (double*)calloc((size_t)(2 * n * N), sizeof(double));
    double* ug = (double*)calloc((size_t)(9 * N), sizeof(double));
    for (int64_t p = 0; p < N; ++p) {
        x[2 * p] = 0.28209479177387814;                 /* n_0^0 = 1/sqrt(4 pi), node-contiguous rows */
        ug[(0 + 3 * 0) * N + p] = 0.5;                  /* ugrad(p,i,k) at plane i + 3k */
        ug[(1 + 3 * 1) * N + p] = 0.5;
        ug[(2 + 3 * 2) * N + p] = -1.0;
    }
    sfb_step_opts o;
    memset(&o, 0, sizeof o);
    o.dt = 0.01; o.iota = 1.0; o.nu_mult = 1.0;
    o.terms = SFB_LROT | SFB_REG; o.scheme = SFB_EULER; o.nsteps = 5;
    if (argc > 2 && strcmp(argv[2], "multi") == 0) {
        int devs[64];
        int nd = sfb_device_count();
        if (nd > 64) nd = 64;
        for (int i = 0; i < nd; ++i) devs[i] = i;
        rc = sfb_step_arr_multi(x, y, N, N, ug, NULL, &o, devs, nd);
    } else {
        rc = sfb_step_arr(x, y, N, N, ug, NULL, &o);
    }
    printf("step %d\n", rc);
    if (rc != SFB_OK) { printf("error %s\n", sfb_last_error()); return 1; }
    for (int j = 0; j < n; ++j) printf("%a %a\n", y[2 * ((int64_t)j * N)], y[2 * ((int64_t)j * N) + 1]);
    sfb_finalize();
    free(x); free(y); free(ug);
    return 0;
}
