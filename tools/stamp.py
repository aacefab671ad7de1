"""Hash of the kernel sources an ncu capture refers to.  profiles/fp64_ops.json and profiles/traffic.json carry the stamp
of the tree they were measured on; bench.py ignores them (and says so in the JSON line) when the tree has moved on, so that
a kernel edit can never silently falsify roofline.frac."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tree_hash():
    h = hashlib.sha1()
    pkg = os.path.join(ROOT, "specfab_b200")
    files = [os.path.join(pkg, "build.py")]
    for sub in ("csrc", "codegen"):
        d = os.path.join(pkg, sub)
        files += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh", ".py"))]
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(tree_hash())
