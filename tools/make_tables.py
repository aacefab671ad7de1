#!/usr/bin/env python3
"""Build specfab_b200/data/gaunt_L20.npz from the reference's Gaunt table (DATA, not source).

The reference fills four real(8) arrays GC, GCm, GC_m1, GC_p1 of shape (231,231,15) with ~43k
assignment statements `GC(i,j,k) = 0.282095` (src/include/gaunt__body.f90, declared in
src/include/gaunt__head.f90:2).  The right-hand sides are un-suffixed Fortran literals, i.e.
real(4): the value the reference actually uses is double(float32(literal)).  This script parses
those statements and stores them as COO (i,j,k 0-based, float32 value) -- float32 storage is
lossless by construction.  Run here (needs /root/reference); the .npz is committed.

Usage: python tools/make_tables.py [/root/reference]
"""
import re, sys, os
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = os.path.join(ref, "src", "include", "gaunt__body.f90")
pat = re.compile(r"^\s*(GC|GCm|GC_m1|GC_p1)\((\d+),(\d+),(\d+)\)\s*=\s*([-+0-9.eE]+)\s*$")
out = {k: ([], [], [], []) for k in ("GC", "GCm", "GC_m1", "GC_p1")}
nline = 0
with open(src) as f:
    for line in f:
        if not line.strip():
            continue
        m = pat.match(line)
        if m is None:
            raise SystemExit("unparsed line: %r" % line)
        name, i, j, k, v = m.groups()
        o = out[name]
        o[0].append(int(i) - 1); o[1].append(int(j) - 1); o[2].append(int(k) - 1)
        o[3].append(np.float32(v))      # real(4) literal semantics
        nline += 1
arrs = {}
for name, (i, j, k, v) in out.items():
    arrs[name + "_i"] = np.asarray(i, np.int16)
    arrs[name + "_j"] = np.asarray(j, np.int16)
    arrs[name + "_k"] = np.asarray(k, np.int8)
    arrs[name + "_v"] = np.asarray(v, np.float32)
    print(name, len(i), "entries")
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "specfab_b200", "data", "gaunt_L20.npz")
np.savez_compressed(dst, Lmax=np.int32(20), nlm_max=np.int32(231), ncat=np.int32(15), **arrs)
print("wrote", os.path.normpath(dst), nline, "statements")
