"""Host-pointer path of sfb_step_arr: pageable numpy arrays vs pinned buffers vs the same arrays page-locked in place."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import specfab_b200 as sf
L, N = 8, 1_000_000
lm, n = sf.init(L)
rng = np.random.default_rng(0)
u = rng.standard_normal((N, 3, 3)); u -= np.eye(3)[None] * (np.trace(u, axis1=1, axis2=2) / 3)[:, None, None]
x = np.zeros((N, n), dtype=np.complex128, order="F"); x[:, 0] = 0.28209479177387814
ug = np.asfortranarray(u)
out = np.empty((N, n), dtype=np.complex128, order="F")
kw = dict(dt=1e-3, terms=("lrot", "reg"), scheme="rk4")
for _ in range(2):
    sf.step_arr(x, ug, out=out, **kw)
t0 = time.perf_counter()
for _ in range(3):
    sf.step_arr(x, ug, out=out, **kw)
print("pageable numpy arrays: %.1f ms per step, %.2e node-updates/s" % ((time.perf_counter() - t0) / 3 * 1e3, N * 3 / (time.perf_counter() - t0)))
def pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype).pin_memory().numpy()
xp = pinned((n, N), torch.complex128).T; xp[:] = x
up = pinned((3, 3, N), torch.float64).transpose(2, 1, 0); up[:] = u
op = pinned((n, N), torch.complex128).T
for _ in range(2):
    sf.step_arr(xp, up, out=op, **kw)
t0 = time.perf_counter()
for _ in range(3):
    sf.step_arr(xp, up, out=op, **kw)
print("pinned arrays: %.1f ms per step, %.2e node-updates/s" % ((time.perf_counter() - t0) / 3 * 1e3, N * 3 / (time.perf_counter() - t0)))
for a in (x, ug, out):
    sf.pin_array(a)
for _ in range(2):
    sf.step_arr(x, ug, out=out, **kw)
t0 = time.perf_counter()
for _ in range(3):
    sf.step_arr(x, ug, out=out, **kw)
print("the same numpy arrays after pin_array (cudaHostRegister): %.1f ms per step, %.2e node-updates/s" % ((time.perf_counter() - t0) / 3 * 1e3, N * 3 / (time.perf_counter() - t0)))
assert np.array_equal(out, op)
for a in (x, ug, out):
    sf.unpin_array(a)
