"""Shared helpers of the parity tests: seeded synthetic inputs and error norms."""
import numpy as np


def lm_list(L):
    return [(l, m) for l in range(0, L + 1, 2) for m in range(-l, l + 1)]


def random_states(L, N, seed, physical=True, decay=0.6):
    """Random spectral states: n00 = 1/sqrt(4 pi), amplitudes decaying with l.  physical=True
    imposes the real-ODF symmetry n_l^-m = (-1)^m conj(n_l^m); False gives general complex
    vectors (the reference operators are defined for those too)."""
    rng = np.random.default_rng(seed)
    lm = lm_list(L)
    n = len(lm)
    x = np.zeros((N, n), dtype=np.complex128)
    for j, (l, m) in enumerate(lm):
        amp = 0.25 * decay ** (l / 2)
        x[:, j] = amp * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    if physical:
        idx = {k: j for j, k in enumerate(lm)}
        for j, (l, m) in enumerate(lm):
            if m == 0:
                x[:, j] = x[:, j].real
            elif m < 0:
                x[:, j] = (-1) ** m * np.conj(x[:, idx[(l, -m)]])
    x[:, 0] = 1 / np.sqrt(4 * np.pi) + (0 if physical else 0.01j * rng.standard_normal(N))
    return x


def random_ugrad(N, seed):
    """standard-normal 3x3, traceless, scaled to ||D||_F = sqrt(1.5)  """
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((N, 3, 3))
    u -= np.eye(3)[None] * (np.trace(u, axis1=1, axis2=2) / 3)[:, None, None]
    D = (u + u.transpose(0, 2, 1)) / 2
    nrm = np.sqrt((D ** 2).sum(axis=(1, 2)))
    return u * (np.sqrt(1.5) / nrm)[:, None, None]


def random_tau(N, seed):
    """independent traceless symmetric normal 3x3 scaled to ||tau||_F = 1"""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((N, 3, 3))
    t = (a + a.transpose(0, 2, 1)) / 2
    t -= np.eye(3)[None] * (np.trace(t, axis1=1, axis2=2) / 3)[:, None, None]
    return t / np.sqrt((t ** 2).sum(axis=(1, 2)))[:, None, None]


def relerr_nodes(a, b):
    """norm-relative error per node: ||a-b||_inf / ||b||_inf   (SURVEY.md 8c tolerances)"""
    a = np.asarray(a).reshape(a.shape[0], -1)
    b = np.asarray(b).reshape(b.shape[0], -1)
    return np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)
