#!/usr/bin/env python3
"""Symbolic (bilinear) mode for tools/f90eval.py.

The orthotropic moment bodies of the reference (src/include/ev_v2/ev_v4/ev_c2b2/ev_c2v2__body.f90, 80-200 KB of
Mathematica-exported text each) are bilinear forms  ev(..) = REAL( sum_pq C_pq * b_p * n_q )  in the harmonic
coefficients of two distributions.  This module interprets the Fortran text ONCE with polynomial values
(constants keep Fortran kind semantics -- real(4) literals and their products are rounded in single precision
until they meet a complex(8) variable) and returns the coefficient tensors C as data.  Nothing of the reference
text is copied; tools/make_orthotropic_tables.py stores only the numbers.
"""
import f90eval as fe
from f90eval import V, binop as vbinop, neg as vneg


class Poly:
    """sum of monomials: key = tuple of variable names (sorted), value = complex coefficient (double)"""
    __slots__ = ("t",)
    k = "poly"

    def __init__(self, t=None):
        self.t = t or {}

    @staticmethod
    def var(name):
        return Poly({(name,): 1 + 0j})

    @staticmethod
    def const(v):
        return Poly({(): complex(v.v)})


def _as_poly(x):
    return x if isinstance(x, Poly) else Poly.const(x)


def padd(a, b, sign=1):
    out = dict(a.t)
    for k, c in b.t.items():
        out[k] = out.get(k, 0j) + sign * c
    return Poly(out)


def pmul(a, b):
    out = {}
    for ka, ca in a.t.items():
        for kb, cb in b.t.items():
            k = tuple(sorted(ka + kb))
            out[k] = out.get(k, 0j) + fe._cmul(complex(ca), complex(cb))
    return Poly(out)


_orig_binop = fe.binop
_orig_neg = fe.neg


def sbinop(op, a, b):
    if not isinstance(a, Poly) and not isinstance(b, Poly):
        return _orig_binop(op, a, b)
    if op == "**":
        if isinstance(b, Poly):
            raise TypeError("variable exponent")
        n = int(b.v)
        if float(b.v) != n or n < 0:
            raise TypeError("non-integer power of a variable")
        r = Poly.const(V("i", 1))
        for _ in range(n):
            r = pmul(r, a)
        return r
    A, B = _as_poly(a), _as_poly(b)
    if op == "+":
        return padd(A, B)
    if op == "-":
        return padd(A, B, -1)
    if op == "*":
        return pmul(A, B)
    if op == "/":
        if isinstance(b, Poly):
            raise TypeError("division by a variable")
        inv = 1.0 / complex(b.v).real if complex(b.v).imag == 0 else 1 / complex(b.v)
        # division by a constant: divide every coefficient (true division, like the run-time expression)
        return Poly({k: complex(c.real / complex(b.v).real, c.imag / complex(b.v).real) if complex(b.v).imag == 0 else c / complex(b.v)
                     for k, c in A.t.items()})
    raise NotImplementedError(op)


def sneg(a):
    if isinstance(a, Poly):
        return Poly({k: -c for k, c in a.t.items()})
    return _orig_neg(a)


def _s_real(a):
    if isinstance(a, Poly):
        return RealOf(a)
    return fe._f_real(a)


class RealOf(Poly):
    """REAL(poly): kept symbolic -- the real part is taken when the form is evaluated"""
    __slots__ = ()

    def __init__(self, p):
        super().__init__(p.t)


def _s_conjg(a):
    """conjugate of a polynomial in REAL variables: conjugate the coefficients"""
    if isinstance(a, Poly):
        return Poly({k: complex(c).conjugate() for k, c in a.t.items()})
    return fe._f_conjg(a)


def install():
    fe.binop = sbinop
    fe.neg = sneg
    fe.INTRINSICS = dict(fe.INTRINSICS, real=_s_real, conjg=_s_conjg)


def uninstall():
    fe.binop = _orig_binop
    fe.neg = _orig_neg
    fe.INTRINSICS = dict(fe.INTRINSICS, real=fe._f_real, conjg=fe._f_conjg)
