"""SASS digest of the compiled step / field kernels: per kernel the instruction counts that the design claims rest on
(DFMA / DMUL / DADD, UMOV immediates, LDCU / LDC constant loads, LDS / STS, LDG / STG, UBLKCP = the 1-D TMA bulk copies,
SYNCS = mbarrier, LDL / STL spills).  usage: python tools/sass_digest.py [tags...] > profiles/r02_sass_digest.json
Default tags: the kernels behind the BASELINE configs (registry variants 0 / 100) and the field kernels."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "specfab_b200", "_build")
OPS = ("DFMA", "DMUL", "DADD", "UMOV", "LDCU", "LDC", "LDS", "STS", "LDG", "STG", "LD", "ST", "UBLKCP", "SYNCS", "BAR", "LDL", "STL",
       "IMAD", "MOV", "BRA", "CALL", "SHFL", "DMMA", "MUFU")


def digest(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = {}
    for fn in re.split(r"Function : ", out)[1:]:
        name = fn.split("\n")[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        short = dem.split("(")[0].replace("void ", "")
        c = collections.Counter()
        n = 0
        for ln in fn.splitlines():
            mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if mm:
                n += 1
                c[mm.group(1)] += 1
        d = {"instructions": n, "code_bytes": 16 * n}
        for op in OPS:
            if c[op]:
                d[op] = c[op]
        key = short
        k = 2
        while key in res:
            key = "%s#%d" % (short, k); k += 1
        res[key] = d
    return res


if __name__ == "__main__":
    tags = sys.argv[1:] or ["step_L8_lrot", "step_L8_lrot_v100", "step_L8_ddrx", "step_L12_ddrx", "step_L20_ddrx", "sfb_fields"]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stamp
    rep = {"_stamp": stamp.tree_hash(), "_how": "cuobjdump -sass of specfab_b200/_build/<tag>.o, opcode histogram per kernel (tools/sass_digest.py)"}
    for t in tags:
        p = os.path.join(OBJ, t + ".o")
        if os.path.exists(p):
            rep[t] = digest(p)
    print(json.dumps(rep, indent=1))
