"""Hash of the kernel sources an ncu capture refers to.  profiles/fp64_ops.json and profiles/traffic.json carry the stamp
of the tree they were measured on; bench.py ignores them (and says so in the JSON line) when the tree has moved on, so that
a kernel edit can never silently falsify roofline.frac."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# sources that only the opt-in experimental variants (SFB_EXP_VARIANTS=1) compile: not part of any default kernel
EXPERIMENTAL = ("sfb_step_wloop.cuh", "emit_wloop.py")


def tree_hash():
    """Everything that determines the DEFAULT kernels: the CUDA sources and code generators (minus the experimental windowed
    kernel) and, from build.py, the variant tables and compiler flags -- not its comments or the experimental variant list."""
    import sys
    h = hashlib.sha1()
    pkg = os.path.join(ROOT, "specfab_b200")
    files = []
    for sub in ("csrc", "codegen"):
        d = os.path.join(pkg, sub)
        files += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh", ".py")) and f not in EXPERIMENTAL]
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    sys.path.insert(0, ROOT)
    from specfab_b200 import build
    for name in ("TUNE", "TUNE_RK", "FULL_DEFAULT", "CFLAGS", "ARCH"):
        v = getattr(build, name)
        v = sorted(v.items()) if isinstance(v, dict) else [x for x in v if not os.path.isabs(x)]      # flags, not this checkout's paths
        h.update(repr(v).encode())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(tree_hash())
