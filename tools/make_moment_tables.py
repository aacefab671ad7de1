#!/usr/bin/env python3
"""Extract the LINEAR coefficient tensors of the reference's 6th/8th-order structure-tensor bodies (DATA, not source).

  a6 = f_ev_c6 <- src/include/ev_c6__body.f90   (src/moments.f90:220-227)    729 entries, harmonics l <= 6
  a8 = f_ev_c8 <- src/include/ev_c8__body.f90   (src/moments.f90:229-236)   6561 entries, harmonics l <= 8

Each body sets k and ev(i1..ik) = REAL( sum_j C_j nlm_j ) with real(4)/complex(4) constants; the caller then scales
by k / f_ev_c0(n00).  tools/f90sym.py interprets the text once with polynomial values (Fortran kind semantics for
the constants).  The reference assigns every one of the 3^k entries separately; this script measures how far index
permutations of one entry differ (they are the same Mathematica expression up to term order) and stores the
canonical (sorted-index) entries: COO (u, j, Re C, Im C) with u the rank of the sorted index tuple in
lexicographic order and j the 0-based position in nlm (0..44).   Output: specfab_b200/data/moments_l8.npz
"""
import itertools, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import f90eval as fe
import f90sym as fs

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
INC = os.path.join(ref, "src", "include")
BODIES = {"c6": ("ev_c6__body.f90", 6), "c8": ("ev_c8__body.f90", 8)}


class AnyKind(dict):
    def __missing__(self, k):
        return "r8"


def pos(name):
    l, m = name[1:].split("_")
    l, m = int(l), int(m)
    return l * (l + 1) // 2 + m


def main():
    fs.install()
    orig_conv = fe._conv
    fe._conv = lambda x, k: x if isinstance(x, fs.Poly) else orig_conv(x, k)
    out = {}
    for tag, (fn, rank) in BODIES.items():
        t0 = time.time()
        env = {"Pi": fe.V("r8", 3.141592653589793), "n00": fs.Poly.var("n0_0")}
        for l in (2, 4, 6, 8):
            env["n%dm" % l] = {m: fs.Poly.var("n%d_%d" % (l, m)) for m in range(-l, l + 1)}
        fe.run_body(open(os.path.join(INC, fn)).read(), env, AnyKind())
        ev = env["ev"]
        assert len(ev) == 3 ** rank, (tag, len(ev))
        uniq = list(itertools.combinations_with_replacement((1, 2, 3), rank))
        rows = {}
        for key, poly in ev.items():
            v = np.zeros(45, dtype=np.complex128)
            for kk, c in poly.t.items():
                assert len(kk) == 1, "not linear: %r" % (kk,)
                v[pos(kk[0])] += c
            rows[key] = v
        dev = 0.0
        U, J, C = [], [], []
        for u, key in enumerate(uniq):
            base = rows[key]
            scale = np.abs(base).max()
            for perm in set(itertools.permutations(key)):
                dev = max(dev, np.abs(rows[perm] - base).max() / scale)
            for j in np.flatnonzero(base):
                U.append(u); J.append(j); C.append(base[j])
        C = np.asarray(C)
        assert np.all((C.real == 0) | (C.imag == 0))
        out[tag + "_u"] = np.asarray(U, np.int16); out[tag + "_j"] = np.asarray(J, np.int8); out[tag + "_c"] = C
        out[tag + "_k"] = np.float64(env["k"].v); out[tag + "_rank"] = np.int64(rank)
        out[tag + "_permdev"] = np.float64(dev)
        print(tag, "rank", rank, "unique", len(uniq), "nnz", len(U), "k", float(env["k"].v), "max permutation deviation", dev,
              "%.1fs" % (time.time() - t0), flush=True)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "moments_l8.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.normpath(dst), os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
