#!/usr/bin/env python3
"""Generate tests/golden/refbodies.npz: golden vectors from the REFERENCE'S OWN SOURCE TEXT.

No Fortran compiler exists in this image, so the reference cannot run.  Instead, the formula
bodies on the hot path are *interpreted* from /root/reference with Fortran kind semantics by
tools/f90eval.py (real(4) literals, left-to-right rounding, REAL()/aimag kinds ...).  The oracle
(oracle/specfab_oracle.py, a hand restatement) must reproduce these outputs -- see
tests/test_oracle_golden.py.  Nothing is copied from the reference; only inputs/outputs are
stored.  Run here:  python tools/make_golden.py [/root/reference]
"""
import os, re, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from f90eval import V, run_body, Parser, tokenize, logical_lines

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRC = os.path.join(ref, "src")
rd = lambda *p: open(os.path.join(SRC, *p)).read()
PI = V("r8", 3.141592653589793)   # src/header.f90:12
rng = np.random.default_rng(20260817)
NCASE = 6


def cvec(vals, lo):
    return {lo + n: V("c8", complex(v)) for n, v in enumerate(vals)}


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        if ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def eval_expr(txt, env):
    return Parser(tokenize(txt), env).expr()


def array_ctor(line, env):
    """evaluate `name = [pre*] [ e1, e2, ... ]` element-wise (optional scalar prefactor)."""
    name, rhs = line.split("=", 1)
    m = re.match(r"\s*(?:(.*?)\*)?\s*\[(.*)\]\s*$", rhs.strip())
    pre, body = m.group(1), m.group(2)
    vals = [eval_expr(e, env) for e in split_top(body)]
    if pre:
        from f90eval import binop
        p = eval_expr(pre, env)
        vals = [binop("*", p, v) for v in vals]
    return name.strip(), vals


out = {}
nlm_in = np.zeros((NCASE, 15), dtype=np.complex128)
for c in range(NCASE):
    v = rng.standard_normal(15) + 1j * rng.standard_normal(15)
    v[0] = 0.28 + 0.05 * rng.standard_normal()      # n00 real-ish but keep a small imaginary part in one case
    if c == NCASE - 1:
        v[0] += 0.01j
    nlm_in[c] = v
out["nlm"] = nlm_in

# ---- ev_c4__body.f90 (a4, real(4) constants)  and ev_c4_Mandel__body.f90 (d0 constants) ----
a4_raw = np.zeros((NCASE, 3, 3, 3, 3)); a4_k = np.zeros(NCASE)
a4m_raw = np.zeros((NCASE, 6, 6)); a4m_k = np.zeros(NCASE)
for c in range(NCASE):
    env = {"Pi": PI, "n00": V("c8", nlm_in[c, 0]), "n2m": cvec(nlm_in[c, 1:6], -2), "n4m": cvec(nlm_in[c, 6:15], -4)}
    run_body(rd("include", "ev_c4__body.f90"), env, {"k": "r8", "ev": "r8"})
    for key, val in env["ev"].items():
        a4_raw[(c,) + tuple(i - 1 for i in key)] = val.v
    a4_k[c] = env["k"].v
    env = {"Pi": PI, "n00": V("c8", nlm_in[c, 0]), "n2m": cvec(nlm_in[c, 1:6], -2), "n4m": cvec(nlm_in[c, 6:15], -4)}
    run_body(rd("include", "ev_c4_Mandel__body.f90"), env, {"k": "r8", "ev": "r8"})
    for key, val in env["ev"].items():
        a4m_raw[(c,) + tuple(i - 1 for i in key)] = val.v
    a4m_k[c] = env["k"].v
out.update(a4_ev=a4_raw, a4_k=a4_k, a4M_ev=a4m_raw, a4M_k=a4m_k)

# ---- ev_c2__body.f90 lines 3-14 (entries before the final affine map) ----
c2_raw = np.zeros((NCASE, 3, 3))
body = rd("include", "ev_c2__body.f90")
stmts = [s for s in logical_lines(body) if re.match(r"\s*ev\(\d,\d\)\s*=", s)]
assert len(stmts) == 9, stmts
import importlib.util
for c in range(NCASE):
    from f90eval import binop
    n00 = V("c8", nlm_in[c, 0])
    hat = {m: binop("/", V("c8", nlm_in[c, 3 + m]), n00) for m in range(3)}   # n2mhat(0:2) = n2m(0:2)/n00
    env = {"n2mhat": hat}
    run_body("\n".join(stmts), env, {"ev": "r8"})
    for key, val in env["ev"].items():
        c2_raw[c, key[0] - 1, key[1] - 1] = val.v
out["c2_ev"] = c2_raw

# ---- ddrx-coupling-weights.f90 (g(1:15), k) and quad_rr / quad_tp / LROT weight vectors ----
dyn = rd("dynamics.f90").splitlines()
def find(pattern, start=0):
    for n in range(start, len(dyn)):
        if re.search(pattern, dyn[n]):
            return n
    raise KeyError(pattern)

n_rr = find(r"function quad_rr")
n_q2m = find(r"q2m = \[", n_rr)
q2m_txt = " ".join(l.split("!")[0].strip().rstrip("&") for l in dyn[n_q2m:n_q2m + 5])
n_tp = find(r"function quad_tp")
q1m_line = dyn[find(r"q1m = fsq1\*\[", n_tp)]
n_lrot = find(r"function M_LROT")
gl = {nm: dyn[find(r"^\s*%s\s*=" % nm, n_lrot)] for nm in
      ("gz_rot", "gn_rot", "gp_rot", "g0_Tay", "gz_Tay", "gn_Tay", "gp_Tay")}

mats = np.zeros((NCASE, 3, 3)); omgs = np.zeros((NCASE, 3, 3))
q2 = np.zeros((NCASE, 5), complex); q1 = np.zeros((NCASE, 3), complex)
gw = np.zeros((NCASE, 4, 6), complex)      # g0, gz, gn, gp of M_LROT for (eps=mats, omg=omgs, iota=1, zeta=0)
gd = np.zeros((NCASE, 15), complex); gd_k = np.zeros(NCASE)
XYZ = {"x": V("i", 1), "y": V("i", 2), "z": V("i", 3)}
for c in range(NCASE):
    A = rng.standard_normal((3, 3)); S = (A + A.T) / 2; S -= np.eye(3) * np.trace(S) / 3
    B = rng.standard_normal((3, 3)); W = (B - B.T) / 2
    mats[c], omgs[c] = S, W
    Menv = {(i + 1, j + 1): V("r8", S[i, j]) for i in range(3) for j in range(3)}
    fsq = V("r8", np.sqrt(2 * 3.141592653589793 / 15))
    env = dict(XYZ, M=Menv, Pi=PI, fsq=fsq, r=V("c8", 1 + 0j), i=V("c8", 1j))
    _, vals = array_ctor(q2m_txt, env)
    q2[c] = [v.v for v in vals]
    Wenv = {(i + 1, j + 1): V("r8", W[i, j]) for i in range(3) for j in range(3)}
    env = dict(XYZ, M=Wenv, Pi=PI, fsq1=V("r8", np.sqrt(2 * 3.141592653589793 / 3)), r=V("c8", 1 + 0j), i=V("c8", 1j))
    _, vals = array_ctor(q1m_line.split("!")[0], env)
    q1[c] = [complex(v.v) for v in vals]
    # LROT weight vectors from the reference's own constructor lines (src/dynamics.f90:78-91)
    env = dict(r=V("c8", 1 + 0j), i=V("c8", 1j), qe=cvec(q2[c], -2), qo=cvec(q1[c], -1))
    parts = {}
    for nm, line in gl.items():
        _, vals = array_ctor(line.split("!")[0], env)
        parts[nm] = np.array([complex(v.v) for v in vals])
    gw[c, 0] = parts["g0_Tay"]
    gw[c, 1] = parts["gz_rot"] + parts["gz_Tay"]
    gw[c, 2] = parts["gn_rot"] + parts["gn_Tay"]
    gw[c, 3] = parts["gp_rot"] + parts["gp_Tay"]
    # DDRX weights with qt = quad_rr(S)
    env = {"Pi": PI, "qt": cvec(q2[c], -2)}
    run_body(rd("include", "ddrx-coupling-weights.f90"), env, {"k": "r8", "g": "c8"})
    gd[c] = [env["g"][n + 1].v for n in range(15)]
    gd_k[c] = env["k"].v
out.update(sym=mats, skew=omgs, quad_rr=q2, quad_tp=q1, lrot_g=gw, ddrx_g=gd, ddrx_k=gd_k)

# ---- orthotropic moment bodies (bilinear in two distributions, real(4) constants): numeric interpretation ----
orth = {}
NO = 3
bq = rng.standard_normal((NO, 15)) + 1j * rng.standard_normal((NO, 15))
nq = rng.standard_normal((NO, 15)) + 1j * rng.standard_normal((NO, 15))
bq[:, 0] = 0.28 + 0.02 * rng.standard_normal(NO); nq[:, 0] = 0.28 + 0.02 * rng.standard_normal(NO)
orth["orth_b"], orth["orth_n"] = bq, nq
for tag, fn, rank in (("v2", "ev_v2__body.f90", 2), ("v4", "ev_v4__body.f90", 4), ("c2b2", "ev_c2b2__body.f90", 4), ("c2v2", "ev_c2v2__body.f90", 4)):
    res = np.zeros((NO,) + (3,) * rank)
    txt = rd("include", fn)
    for c in range(NO):
        env = {"Pi": PI, "b00": V("c8", bq[c, 0]), "b2m": cvec(bq[c, 1:6], -2), "b4m": cvec(bq[c, 6:15], -4),
               "n00": V("c8", nq[c, 0]), "n2m": cvec(nq[c, 1:6], -2), "n4m": cvec(nq[c, 6:15], -4)}
        run_body(txt, env, {"k": "r8", "ev": "r8", "norm": "r8"})
        for key, val in env["ev"].items():
            res[(c,) + tuple(i - 1 for i in key)] = val.v * env["k"].v / env["norm"].v      # ev = ev * k/norm  (src/moments.f90:257)
    orth["orth_" + tag] = res
out.update(orth)

# ---- 6th / 8th order structure tensors (linear in nlm, l <= 8, real(4) constants): numeric interpretation ----
NH = 2
hq = rng.standard_normal((NH, 45)) + 1j * rng.standard_normal((NH, 45))
hq[:, 0] = 0.28 + 0.02 * rng.standard_normal(NH)
out["hi_nlm"] = hq
for tag, fn, rank in (("c6", "ev_c6__body.f90", 6), ("c8", "ev_c8__body.f90", 8)):
    res = np.zeros((NH,) + (3,) * rank)
    txt = rd("include", fn)
    for c in range(NH):
        env = {"Pi": PI, "n00": V("c8", hq[c, 0]), "n2m": cvec(hq[c, 1:6], -2), "n4m": cvec(hq[c, 6:15], -4),
               "n6m": cvec(hq[c, 15:28], -6), "n8m": cvec(hq[c, 28:45], -8)}
        run_body(txt, env, {"k": "r8", "ev": "r8"})
        c0 = float(np.sqrt(4 * np.pi) * hq[c, 0].real)           # f_ev_c0 = REAL(sqrt(4*Pi)*n00)  (src/moments.f90:184-189)
        for key, val in env["ev"].items():
            res[(c,) + tuple(i - 1 for i in key)] = val.v * env["k"].v / c0      # ev = ev * k/f_ev_c0(n00)  (:226,:235)
    out["hi_" + tag] = res

# ---- state ingest a2/a4/a6 -> nlm (affine maps, real(4) constants, conjg rows): numeric interpretation ----
# arbitrary (non-symmetric) arrays: the bodies read specific entries, which pins the index handling
import itertools
NI = 2
SQ2 = float(np.sqrt(2.0))
for tag, fn, rank, nrow in (("a2", "a2_to_nlm__body.f90", 2, 6), ("a4", "a4_to_nlm__body.f90", 4, 15), ("a6", "a6_to_nlm__body.f90", 6, 28)):
    A = rng.standard_normal((NI,) + (3,) * rank)
    res = np.zeros((NI, nrow), dtype=np.complex128)
    for c in range(NI):
        env = {"Pi": PI}
        if tag == "a4":
            pairs = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
            man = {}
            for i in range(6):          # a4_to_mat, src/mandel.f90:52-66
                for j in range(6):
                    f = 2.0 if (i >= 3 and j >= 3) else (SQ2 if (i >= 3 or j >= 3) else 1.0)
                    man[(i + 1, j + 1)] = V("r8", f * A[c][pairs[i] + pairs[j]])
            env["a4Mandel"] = man
        else:
            env[tag] = {tuple(i + 1 for i in idx): V("r8", A[c][idx]) for idx in itertools.product(range(3), repeat=rank)}
        run_body(rd("include", fn), env, {"nlm": "c8"})
        res[c] = [env["nlm"][k + 1].v for k in range(nrow)]
    out["ingest_" + tag] = A
    out["ingest_" + tag + "_nlm"] = res

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "refbodies.npz")
np.savez_compressed(dst, **out)
print("wrote", os.path.normpath(dst), {k: v.shape for k, v in out.items()})
