#!/usr/bin/env python3
"""Extract the bilinear coefficient tensors of the reference's orthotropic moment bodies (DATA, not source).

  a2_orth       <- src/include/ev_v2__body.f90     (src/moments.f90:242-258)
  a4_orth       <- src/include/ev_v4__body.f90     (src/moments.f90:260-276)
  a4_joint      <- src/include/ev_c2b2__body.f90   (src/moments.f90:278-293)
  a4_jointcross <- src/include/ev_c2v2__body.f90   (src/moments.f90:295-311)

Each body sets k (constant), norm = REAL(bilinear form) and ev(...) = REAL(bilinear form in b_p, n_q), with
real(4)/complex(4) constants.  tools/f90sym.py interprets the text once with polynomial values (Fortran kind
semantics for the constants) and this script stores, per body, COO lists (entry, p, q, Re C, Im C) with p, q the
0-based positions of b_p / n_q in the nlm vector (0..14).  Output: specfab_b200/data/orthotropic_l4.npz
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import f90eval as fe
import f90sym as fs

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
INC = os.path.join(ref, "src", "include")
BODIES = {"v2": "ev_v2__body.f90", "v4": "ev_v4__body.f90", "c2b2": "ev_c2b2__body.f90", "c2v2": "ev_c2v2__body.f90"}


def pos(name):
    """'b2_-1' -> (which, index in nlm)"""
    which, rest = name[0], name[1:]
    l, m = rest.split("_")
    l, m = int(l), int(m)
    return which, l * (l + 1) // 2 + m


class AnyKind(dict):
    def __missing__(self, k):
        return "r8"


def sym(prefix):
    d = {"%s00" % prefix: fs.Poly.var("%s0_0" % prefix)}
    d["%s2m" % prefix] = {m: fs.Poly.var("%s2_%d" % (prefix, m)) for m in range(-2, 3)}
    d["%s4m" % prefix] = {m: fs.Poly.var("%s4_%d" % (prefix, m)) for m in range(-4, 5)}
    d["%s6m" % prefix] = {m: fs.Poly.const(fe.V("i", 0)) for m in range(-6, 7)}
    return d


def coo(poly):
    P, Q, C = [], [], []
    for key, c in poly.t.items():
        if c == 0:
            continue
        assert len(key) == 2, "not bilinear: %r" % (key,)
        a, b = pos(key[0]), pos(key[1])
        (wb, p), (wn, q) = (a, b) if a[0] == "b" else (b, a)
        assert wb == "b" and wn == "n", key
        P.append(p); Q.append(q); C.append(c)
    return P, Q, C


def main():
    fs.install()
    orig_conv = fe._conv
    fe._conv = lambda x, k: x if isinstance(x, fs.Poly) else orig_conv(x, k)
    out = {}
    for tag, fn in BODIES.items():
        env = {"Pi": fe.V("r8", 3.141592653589793)}
        env.update(sym("b")); env.update(sym("n"))
        fe.run_body(open(os.path.join(INC, fn)).read(), env, AnyKind())
        ev = env["ev"]
        rank = len(next(iter(ev)))
        E, P, Q, C = [], [], [], []
        for key, poly in ev.items():
            e = sum((i - 1) * 3 ** d for d, i in enumerate(key))      # Fortran column-major position
            p, q, c = coo(poly)
            E += [e] * len(p); P += p; Q += q; C += c
        assert len(ev) == 3 ** rank, (tag, len(ev))
        npq = coo(env["norm"])
        out[tag + "_e"] = np.asarray(E, np.int16); out[tag + "_p"] = np.asarray(P, np.int8); out[tag + "_q"] = np.asarray(Q, np.int8)
        out[tag + "_c"] = np.asarray(C, np.complex128)
        out[tag + "_norm_p"] = np.asarray(npq[0], np.int8); out[tag + "_norm_q"] = np.asarray(npq[1], np.int8)
        out[tag + "_norm_c"] = np.asarray(npq[2], np.complex128)
        out[tag + "_k"] = np.float64(env["k"].v)
        out[tag + "_rank"] = np.int32(rank)
        print(tag, "rank", rank, "nnz", len(E), "norm terms", len(npq[0]), "k", float(env["k"].v))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "orthotropic_l4.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
