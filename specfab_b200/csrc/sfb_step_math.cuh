// Device math shared by every step-kernel skeleton: complex helpers, the quadric coefficients quad_rr / quad_tp
// (src/dynamics.f90:563-593) and the raw DDRX coupling weights (src/include/ddrx-coupling-weights.f90:1-16).
// Included INSIDE the anonymous namespace of a step translation unit (needs SFB_DDRX and sfb_common.cuh).
// ---- complex helpers (forcing preparation) ----
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 cneg(double2 a) { return make_double2(-a.x, -a.y); }

// src/dynamics.f90:563-579 ; q[0..4] <-> m = -2..2 ; M symmetric, row-major m[3][3]
__device__ __forceinline__ void quad_rr(const double m[3][3], double2 q[5]) {
    const double fsq = 0x1.4b5eee37a973cp-1;    // sqrt(2*Pi/15)
    const double sp5 = 0x1.95d83f429fefap-1;    // sqrt(Pi/5)
    const double xx = m[0][0], yy = m[1][1], zz = m[2][2], xy = m[0][1], xz = m[0][2], yz = m[1][2];
    q[0] = make_double2(fsq * (xx - yy), fsq * (2 * xy));
    q[1] = make_double2((2 * fsq) * xz, (2 * fsq) * yz);
    q[2] = make_double2(-((SFB_TWOTHIRDS_F * sp5) * (xx + yy - 2 * zz)), 0.0);
    q[3] = make_double2(-((2 * fsq) * xz), (2 * fsq) * yz);
    q[4] = make_double2(fsq * (xx - yy), fsq * (-2 * xy));
}

// src/dynamics.f90:581-593 ; q[0..2] <-> m = -1..1 ; M antisymmetric
__device__ __forceinline__ void quad_tp(const double m[3][3], double2 q[3]) {
    const double fsq1 = 0x1.727bdd17583bbp+0;   // sqrt(2*Pi/3)
    const double xy = m[0][1], xz = m[0][2], yz = m[1][2];
    q[0] = make_double2(fsq1 * yz, fsq1 * (-xz));
    q[1] = make_double2(fsq1 * (SFB_SQRT2_F * xy), 0.0);
    q[2] = make_double2(fsq1 * (-yz), fsq1 * (-xz));
}

#if SFB_DDRX
// src/include/ddrx-coupling-weights.f90:1-16 with real(4) constants; qt**(2.0) == qt*qt
__device__ __forceinline__ void ddrx_weights_raw(const double2 qt[5], double2 g[15]) {
    const double s5 = 0x1.1e377ap+1, s15 = 0x1.3988e2p+0, s6 = 0x1.3988e2p+1, s2 = 0x1.6a09e6p+0, s3 = 0x1.bb67aep+0;
    const double c2s14 = 0x1.deeea2p+2;   // 2*Sqrt((14.0))   real(4)
    const double c4s7 = 0x1.52a7fap+3;    // 4*Sqrt((7.0))    real(4)
    const double c3s5 = 0x1.ad5338p+2;    // 3.*Sqrt((5.0))   real(4)
    const double2 qm2 = qt[0], qm1 = qt[1], q0 = qt[2], qp1 = qt[3], qp2 = qt[4];
    const double2 q0q0 = cmul(q0, q0), qm1qm1 = cmul(qm1, qm1), qp1qp1 = cmul(qp1, qp1);
    double2 t;
    t = cadd(cadd(q0q0, cmul(cscale(-2.0, qm1), qp1)), cmul(cscale(2.0, qm2), qp2));
    t = cscale(7.0, t); g[0] = make_double2(t.x / s5, t.y / s5);
    g[1] = cadd(cscale(s15, qm1qm1), cmul(cscale(-2.0, q0), qm2));
    g[2] = cadd(cmul(q0, qm1), cmul(cscale(-s6, qp1), qm2));
    g[3] = cadd(cadd(q0q0, cmul(cscale(-1.0, qm1), qp1)), cmul(cscale(-2.0, qm2), qp2));
    g[4] = cadd(cmul(q0, qp1), cmul(cscale(-s6, qm1), qp2));
    g[5] = cadd(cscale(s15, qp1qp1), cmul(cscale(-2.0, q0), qp2));
    t = cneg(cscale(c2s14, cmul(qm2, qm2))); g[6] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cmul(cscale(c4s7, qm1), qm2)); g[7] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cscale(s2, qm1qm1), cmul(cscale(s3, q0), qm2)))); g[8] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cmul(cscale(s6, q0), qm1), cmul(qp1, qm2)))); g[9] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cadd(cscale(3.0, q0q0), cmul(cscale(4.0, qm1), qp1)), cmul(qm2, qp2))));
    g[10] = make_double2(t.x / c3s5, t.y / c3s5);
    t = cneg(cscale(4.0, cadd(cmul(cscale(s6, q0), qp1), cmul(qm1, qp2)))); g[11] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(4.0, cadd(cscale(s2, qp1qp1), cmul(cscale(s3, q0), qp2)))); g[12] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cmul(cscale(c4s7, qp1), qp2)); g[13] = make_double2(t.x / 3.0, t.y / 3.0);
    t = cneg(cscale(c2s14, cmul(qp2, qp2))); g[14] = make_double2(t.x / 3.0, t.y / 3.0);
}
#endif

// catalyst index k of (lk,mk) and of its mirror (lk,-mk); k = 0 | 1..5 | 6..14
__device__ __forceinline__ int cat_mirror(int k) { return k == 0 ? 0 : (k < 6 ? 6 - k + 0 : 20 - k); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }


// L2 prefetch of the mirror rows (m < 0) of a tile, issued next to the bulk copies of the rows m >= 0: the real-ODF test reads them
// from global memory AFTER the tile has landed (the registers are needed by the forcing preparation in between), and would
// otherwise wait for DRAM at that point.  Lanes lane0, lane0 + nlanes, ... each take full-form rows j = hrow(l) + m with m < 0.
__device__ __forceinline__ void prefetch_mirror_rows(const double2* nlm_in, long long ld, long long node0, int nvalid, int L, int lane0, int nlanes) {
    const int ncoef = (L + 1) * (L + 2) / 2;
    const unsigned bytes = (unsigned)nvalid * 16u;
    for (int j = lane0; j < ncoef; j += nlanes) {
        int l = 0;
        while ((l + 2) * (l + 1) / 2 <= j) l += 2;           // first row of degree l + 2 is (l + 2)(l + 1)/2
        const int m = j - l * (l + 1) / 2;
        if (m < 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nlm_in + (long long)j * ld + node0), "r"(bytes) : "memory");
    }
}
