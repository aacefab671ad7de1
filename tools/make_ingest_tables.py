#!/usr/bin/env python3
"""Extract the affine maps a2 -> nlm(1:6), a4 -> nlm(1:15), a6 -> nlm(1:28) of the reference (DATA, not source).

  a2_to_nlm <- src/include/a2_to_nlm__body.f90   (src/moments.f90:68-74)
  a4_to_nlm <- src/include/a4_to_nlm__body.f90   (src/moments.f90:76-84; reads a4Mandel = a4_to_mat(a4), src/mandel.f90:52-66)
  a6_to_nlm <- src/include/a6_to_nlm__body.f90   (src/moments.f90:86-92)

Every nlm(k) is a constant plus a complex-weighted sum of REAL tensor entries (real(4) literals; the m < 0 rows are
conjg() of the m > 0 rows).  tools/f90sym.py interprets the text once; the a4 map is composed with a4_to_mat so
that all three tables address the tensor itself:  nlm[row] = c0[row] + sum_t C_t * A[flat_t], flat = Fortran
column-major position in the 3^k array.     Output: specfab_b200/data/ingest_l6.npz
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import f90eval as fe
import f90sym as fs

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
INC = os.path.join(ref, "src", "include")
SQRT2 = float(np.sqrt(2.0))     # s = sqrt(2.0d0), src/mandel.f90


def mandel_of_a4():
    """a4Mandel(i,j) (1-based) -> (factor, 4-index) exactly as src/mandel.f90:60-65"""
    pairs = [(1, 1), (2, 2), (3, 3), (2, 3), (1, 3), (1, 2)]
    out = {}
    for i in range(6):
        for j in range(6):
            f = (SQRT2 if i >= 3 else 1.0) * (SQRT2 if j >= 3 else 1.0)
            if i >= 3 and j >= 3:
                f = 2.0
            out[(i + 1, j + 1)] = (f, pairs[i] + pairs[j])
    return out


class AnyKind(dict):
    def __missing__(self, k):
        return "c8"


def flat(idx):
    return sum((i - 1) * 3 ** d for d, i in enumerate(idx))


def main():
    fs.install()
    orig_conv = fe._conv
    fe._conv = lambda x, k: x if isinstance(x, fs.Poly) else orig_conv(x, k)
    out = {}
    man = mandel_of_a4()
    for tag, fn, rank, nrow in (("a2", "a2_to_nlm__body.f90", 2, 6), ("a4", "a4_to_nlm__body.f90", 4, 15), ("a6", "a6_to_nlm__body.f90", 6, 28)):
        env = {"Pi": fe.V("r8", 3.141592653589793)}
        if tag == "a4":
            env["a4Mandel"] = {k: fs.Poly.var("M%d_%d" % k) for k in man}
        else:
            import itertools
            env[tag] = {idx: fs.Poly.var("A" + "_".join(map(str, idx))) for idx in itertools.product((1, 2, 3), repeat=rank)}
        fe.run_body(open(os.path.join(INC, fn)).read(), env, AnyKind())
        nlm = env["nlm"]
        assert sorted(nlm) == list(range(1, nrow + 1)), sorted(nlm)
        R, F, C = [], [], []
        c0 = np.zeros(nrow, dtype=np.complex128)
        for row in range(1, nrow + 1):
            p = nlm[row]
            p = p if isinstance(p, fs.Poly) else fs.Poly.const(p)
            acc = {}
            for key, c in p.t.items():
                if len(key) == 0:
                    c0[row - 1] += c
                    continue
                assert len(key) == 1, key
                name = key[0]
                if name[0] == "M":
                    f, idx = man[tuple(int(x) for x in name[1:].split("_"))]
                    acc[flat(idx)] = acc.get(flat(idx), 0j) + c * f
                else:
                    idx = tuple(int(x) for x in name[1:].split("_"))
                    acc[flat(idx)] = acc.get(flat(idx), 0j) + c
            for f_, c in sorted(acc.items()):
                if c != 0:
                    R.append(row - 1); F.append(f_); C.append(c)
        out[tag + "_row"] = np.asarray(R, np.int16); out[tag + "_flat"] = np.asarray(F, np.int16)
        out[tag + "_c"] = np.asarray(C, np.complex128); out[tag + "_c0"] = c0
        print(tag, "rows", nrow, "nnz", len(R), "c0[0]", c0[0])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "ingest_l6.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.normpath(dst), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
