#!/usr/bin/env python3
"""A tiny Fortran-90 *expression* interpreter with Fortran kind semantics.

Purpose (test infrastructure only): the reference cannot be compiled in this image (no Fortran
compiler), so the oracle in oracle/ is a hand restatement.  To pin the riskiest part of that
restatement -- the machine-generated formula bodies with their default-real (real(4)) literals --
this module evaluates the reference's own source TEXT (read from /root/reference at generation
time, never copied) the way gfortran would:

  * un-suffixed real literals (`3.`, `(5.0)`, `1.5`) are real(4); `d0` literals are real(8);
  * integer (op) real(4) -> real(4); anything (op) real(8) -> real(8); `(0,1)` is complex(4);
  * every operation is rounded in the kind of its promoted operands (numpy float32/complex64
    arithmetic for kind 4), evaluated left to right with Fortran precedence
    (`**` > `* /` > unary sign > `+ -`);
  * intrinsics of real(4) arguments return real(4) (`Sqrt((5.0))` -> float32);
  * `REAL(z)` / `aimag(z)` keep the kind of z; complex division follows GCC's
    -fcx-fortran-rules expansion (Smith's range reduction, true divisions);
  * `z**(2.0)` with complex z is evaluated as z*z (gfortran lowers it to cpow, which agrees
    with z*z to a few ulp; documented in DESIGN.md).

Only what the hot-path include bodies need is implemented (assignments of scalar expressions,
indexed names, `&` continuations, `!` comments).  Used by tools/make_golden.py.
"""
import math
import re
import numpy as np

_KRANK = {"i": 0, "r4": 1, "r8": 2, "c4": 3, "c8": 4}


class V:
    """A typed Fortran scalar."""
    __slots__ = ("k", "v")

    def __init__(self, k, v):
        self.k = k
        if k == "i":
            self.v = int(v)
        elif k == "r4":
            self.v = np.float32(v)
        elif k == "r8":
            self.v = float(v)
        elif k == "c4":
            self.v = np.complex64(v)
        else:
            self.v = complex(v)

    def __repr__(self):
        return "V(%s,%r)" % (self.k, self.v)


def _promote(a, b):
    ka, kb = a.k, b.k
    cplx = ka[0] == "c" or kb[0] == "c"
    dbl = "8" in ka or "8" in kb
    if cplx:
        k = "c8" if dbl else "c4"
    elif ka == "i" and kb == "i":
        k = "i"
    else:
        k = "r8" if dbl else "r4"
    return k


def _conv(x, k):
    if x.k == k:
        return x
    if k in ("r4", "r8", "i"):
        if x.k[0] == "c":
            raise TypeError("complex -> real conversion needs REAL()")
        return V(k, x.v)
    return V(k, x.v)  # real/int -> complex (imag 0) or c4 -> c8


def _cdiv(a, b):
    """GCC expand_complex_division, flag_complex_method=1 (Fortran rules)."""
    ar, ai, br, bi = a.real, a.imag, b.real, b.imag
    if abs(br) < abs(bi):
        ratio = br / bi
        div = (br * ratio) + bi
        tr = (ar * ratio) + ai
        ti = (ai * ratio) - ar
    else:
        ratio = bi / br
        div = (bi * ratio) + br
        tr = (ai * ratio) + ar
        ti = ai - (ar * ratio)
    return complex(tr / div, ti / div)


def _cmul(a, b):
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def binop(op, a, b):
    if op == "**":
        return _pow(a, b)
    k = _promote(a, b)
    x, y = _conv(a, k).v, _conv(b, k).v
    if k == "i":
        if op == "+":
            return V("i", x + y)
        if op == "-":
            return V("i", x - y)
        if op == "*":
            return V("i", x * y)
        q = abs(x) // abs(y)
        return V("i", q if (x >= 0) == (y >= 0) else -q)
    if k == "c8":
        if op == "+":
            return V(k, complex(x.real + y.real, x.imag + y.imag))
        if op == "-":
            return V(k, complex(x.real - y.real, x.imag - y.imag))
        if op == "*":
            return V(k, _cmul(x, y))
        return V(k, _cdiv(x, y))
    with np.errstate(all="ignore"):
        if op == "+":
            return V(k, x + y)
        if op == "-":
            return V(k, x - y)
        if op == "*":
            return V(k, x * y)
        return V(k, x / y)


def _pow(a, b):
    if b.k == "i":
        n = b.v
        if n < 0:
            raise NotImplementedError
        r = V(a.k, 1)
        for _ in range(n):
            r = binop("*", r, a)
        return r
    if a.k[0] == "c":
        if float(b.v) == 2.0:
            return binop("*", a, a)
        raise NotImplementedError("complex ** real only for exponent 2.0")
    k = _promote(a, b)
    x, y = _conv(a, k).v, _conv(b, k).v
    if k == "r4":
        # gfortran folds constant powers with MPFR (correctly rounded); numpy's powf is 1 ulp off for
        # 3.0**1.5, so round the double result instead
        return V("r4", np.float32(math.pow(float(x), float(y))))
    return V("r8", math.pow(x, y))


def neg(a):
    return V(a.k, -a.v)


def _f_sqrt(a):
    if a.k == "i":
        raise TypeError("sqrt(integer)")
    if a.k == "r4":
        return V("r4", np.sqrt(np.float32(a.v)))
    if a.k == "r8":
        return V("r8", math.sqrt(a.v))
    raise NotImplementedError


def _f_real(a):
    if a.k == "c8":
        return V("r8", a.v.real)
    if a.k == "c4":
        return V("r4", a.v.real)
    if a.k == "i":
        return V("r4", a.v)
    return a


def _f_aimag(a):
    if a.k == "c8":
        return V("r8", a.v.imag)
    if a.k == "c4":
        return V("r4", a.v.imag)
    raise TypeError


def _f_conjg(a):
    if a.k[0] != "c":
        raise TypeError("conjg of a non-complex value")
    return V(a.k, a.v.conjugate())


INTRINSICS = {"sqrt": _f_sqrt, "real": _f_real, "aimag": _f_aimag, "conjg": _f_conjg}

_TOK = re.compile(r"\s*(?:(\d+\.\d*(?:[dDeE][-+]?\d+)?|\.\d+(?:[dDeE][-+]?\d+)?|\d+[dDeE][-+]?\d+|\d+)|([A-Za-z_][A-Za-z_0-9]*)|(\*\*|[-+*/(),=]))")


def tokenize(s):
    out, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise SyntaxError("bad token at %r" % s[pos:pos + 20])
        num, name, op = m.groups()
        if num is not None:
            out.append(("num", num))
        elif name is not None:
            out.append(("name", name))
        else:
            out.append(("op", op))
        pos = m.end()
    return out


def _literal(txt):
    t = txt.lower()
    if "d" in t:
        return V("r8", float(t.replace("d", "e")))
    if "." in t or "e" in t:
        return V("r4", np.float32(t))
    return V("i", int(t))


class Parser:
    def __init__(self, toks, env):
        self.t, self.p, self.env = toks, 0, env

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else (None, None)

    def eat(self, kind=None, val=None):
        tk = self.peek()
        if (kind and tk[0] != kind) or (val and tk[1] != val):
            raise SyntaxError("expected %s %s got %s" % (kind, val, tk))
        self.p += 1
        return tk

    # level-2: [sign] term {(+|-) term}
    def expr(self):
        tk = self.peek()
        sign = None
        if tk == ("op", "+") or tk == ("op", "-"):
            sign = self.eat()[1]
        x = self.term()
        if sign == "-":
            x = neg(x)
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.eat()[1]
            y = self.term()
            x = binop(op, x, y)
        return x

    def term(self):
        x = self.factor()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.eat()[1]
            y = self.factor()
            x = binop(op, x, y)
        return x

    def factor(self):
        x = self.primary()
        if self.peek() == ("op", "**"):
            self.eat()
            y = self.factor()  # right associative
            x = binop("**", x, y)
        return x

    def _signed_expr(self):
        return self.expr()

    def primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.eat()
            return _literal(val)
        if kind == "op" and val == "(":
            self.eat()
            a = self.expr()
            if self.peek() == ("op", ","):  # complex literal
                self.eat()
                b = self.expr()
                self.eat("op", ")")
                dbl = a.k == "r8" or b.k == "r8"
                return V("c8" if dbl else "c4", complex(float(a.v), float(b.v)))
            self.eat("op", ")")
            return a
        if kind == "name":
            self.eat()
            args = None
            if self.peek() == ("op", "("):
                self.eat()
                args = [self.expr()]
                while self.peek() == ("op", ","):
                    self.eat()
                    args.append(self.expr())
                self.eat("op", ")")
            low = val.lower()
            if args is not None and low in INTRINSICS and val not in self.env:
                return INTRINSICS[low](*args)
            if val not in self.env:
                raise NameError(val)
            obj = self.env[val]
            if args is None:
                return obj
            key = tuple(a.v for a in args)
            return obj[key if len(key) > 1 else key[0]]
        raise SyntaxError("unexpected %s" % (self.peek(),))


def logical_lines(text):
    """Join `&` continuations, drop comments and blanks."""
    buf = ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:]
        if s.endswith("&"):
            buf += s[:-1] + " "
            continue
        yield buf + s
        buf = ""
    if buf:
        yield buf


def run_body(text, env, lhs_kinds):
    """Execute `name(idx..) = expr` / `name = expr` statements.  lhs_kinds: name -> kind of the
    declared variable (values are converted on assignment, like Fortran)."""
    for stmt in logical_lines(text):
        toks = tokenize(stmt)
        # find the top-level '='
        depth, eq = 0, None
        for n, (k, v) in enumerate(toks):
            if k == "op" and v == "(":
                depth += 1
            elif k == "op" and v == ")":
                depth -= 1
            elif k == "op" and v == "=" and depth == 0:
                eq = n
                break
        if eq is None:
            raise SyntaxError("not an assignment: " + stmt)
        lhs, rhs = toks[:eq], toks[eq + 1:]
        name = lhs[0][1]
        val = Parser(rhs, env).expr()
        kind = lhs_kinds[name]
        if kind[0] != "c" and val.k[0] == "c":
            raise TypeError("complex assigned to real in: " + stmt)
        val = _conv(val, kind)
        if len(lhs) == 1:
            env[name] = val
        else:
            idx = Parser(lhs[2:-1] + [("op", ")")], env)
            keys = [idx.expr()]
            while idx.peek() == ("op", ","):
                idx.eat()
                keys.append(idx.expr())
            key = tuple(a.v for a in keys)
            env.setdefault(name, {})[key if len(key) > 1 else key[0]] = val
    return env
