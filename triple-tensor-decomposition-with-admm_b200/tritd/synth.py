"""Seeded synthetic workloads for the five BASELINE.json configs.

Everything is generated on the host with numpy's PCG64 in column-major order,
so the GPU library, the oracle and the golden fixtures all see the same bytes.
The option sets are the ones the reference's callers pass
(traffic_triple_comparison.m:42-47, video_triple_comparison.m:41-46).
"""
from __future__ import annotations

import numpy as np

TRAFFIC_OPTS = dict(mu=1e-3, rho=1.25, **{"lambda": 1.8}, lambda2=1e-3, maxIter=100, tol=1e-5, disp=0)
VIDEO_OPTS = dict(mu=1e-2, rho=1.2, **{"lambda": 1.8}, lambda2=1e-2, maxIter=100, tol=1e-5, disp=0)

CONFIGS = {
    # name: (n1, n2, n3, r, kind, outlier/missing fraction, seed, opts)
    "cfg1": (50, 50, 50, 5, "lowrank_sparse", 0.10, 1, TRAFFIC_OPTS),
    "cfg2": (256, 256, 200, 5, "traffic", 0.10, 2, TRAFFIC_OPTS),
    "cfg3": (240, 320, 300, 5, "video", 0.0, 3, VIDEO_OPTS),
    "cfg4": (512, 512, 512, 6, "lowrank_sparse", 0.20, 4, TRAFFIC_OPTS),
    "cfg5": (1024, 1024, 512, 8, "lowrank_sparse", 0.10, 5, TRAFFIC_OPTS),
}

DESCRIPTIONS = {
    "cfg1": "synthetic 50x50x50 low-rank + 10% sparse, r=5",
    "cfg2": "traffic-like synthetic 256x256x200, r=5, 10% missing treated as corrupted",
    "cfg3": "video background (Highway-shaped) 240x320x300, r=5",
    "cfg4": "synthetic 512x512x512, r=6, 20% sparse outliers",
    "cfg5": "synthetic 1024x1024x512, r=8, 10% sparse outliers",
}


def cp_r2(A1, B2, C3, t0=0, t1=None):
    """L(i,j,t) = sum_k A1(i,k) B2(j,k) C3(t,k) for t in [t0,t1) (column-major)."""
    C3s = C3[t0:t1]
    n1, n2, n3 = A1.shape[0], B2.shape[0], C3s.shape[0]
    H = (A1[:, None, :] * B2[None, :, :]).reshape(n1 * n2, -1)      # (i,j) row-major
    L = H @ C3s.T                                                   # (i*n2+j, t)
    return np.asfortranarray(L.reshape(n1, n2, n3))


def _truth_factors(n1, n2, n3, r, rng):
    R = r * r
    return rng.standard_normal((n1, R)), rng.standard_normal((n2, R)), rng.standard_normal((n3, R))


def init_factors(n1, n2, n3, r, seed):
    """A0 (n1,r,r), B0 (r,n2,r), C0 (r,r,n3) ~ N(0,1), the injected initialisation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A0 = np.asfortranarray(rng.standard_normal((n1, r, r)))
    B0 = np.asfortranarray(rng.standard_normal((r, n2, r)))
    C0 = np.asfortranarray(rng.standard_normal((r, r, n3)))
    return A0, B0, C0


def make_lowrank_sparse(n1, n2, n3, r, frac, seed, t0=0, t1=None, with_truth=False):
    """D = L0 + S: L0 = CP-r^2 of N(0,1) factors, S has `frac` uniformly placed
    non-zeros ~ U(-10 sigma, 10 sigma), sigma = r (std of an L0 entry).
    Generated slab-wise in t so big tensors never need more than one slab of
    temporaries; slab [t0,t1) of the same seed is identical to slicing the full tensor."""
    t1 = n3 if t1 is None else t1
    rng = np.random.Generator(np.random.PCG64(seed))
    A1, B2, C3 = _truth_factors(n1, n2, n3, r, rng)
    sigma = float(r)
    D = np.empty((n1, n2, t1 - t0), order="F")
    L0 = np.empty_like(D) if with_truth else None
    step = max(1, (1 << 24) // (n1 * n2))
    for a in range((t0 // step) * step, t1, step):        # global blocks, so any slab equals a slice of the whole
        b = min(n3, a + step)
        Ls = cp_r2(A1, B2, C3, a, b)
        srng = np.random.Generator(np.random.PCG64([seed, a]))   # one independent stream per block
        mask = srng.random(Ls.shape) < frac
        S = srng.uniform(-10 * sigma, 10 * sigma, Ls.shape)
        lo, hi = max(a, t0), min(b, t1)
        if with_truth:
            L0[:, :, lo - t0:hi - t0] = Ls[:, :, lo - a:hi - a]
        D[:, :, lo - t0:hi - t0] = (Ls + np.where(mask, S, 0.0))[:, :, lo - a:hi - a]
    return (D, L0) if with_truth else D


def make_traffic(n1, n2, n3, r, frac, seed, with_truth=False):
    """Non-negative counts-like low-rank tensor (mean ~50) with `frac` of the
    entries set to 0 (missing treated as corruption, traffic_triple_comparison.m:29-35)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A1, B2, C3 = _truth_factors(n1, n2, n3, r, rng)
    L0 = cp_r2(np.abs(A1), np.abs(B2), np.abs(C3))
    L0 *= 50.0 / L0.mean()
    D = L0.copy(order="F")
    D[rng.random(D.shape) < frac] = 0.0
    return (D, L0) if with_truth else D


def make_video(n1, n2, n3, seed, with_truth=False):
    """Highway-shaped clip: smoothed static background U(50,200), slow
    illumination drift, a few moving 10x10 bright blocks, N(0,2) noise, clipped to [0,255]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bg = rng.uniform(50, 200, (n1, n2))
    for _ in range(3):  # cheap box smoothing
        bg = (bg + np.roll(bg, 1, 0) + np.roll(bg, -1, 0) + np.roll(bg, 1, 1) + np.roll(bg, -1, 1)) / 5.0
    t = np.arange(n3)
    illum = 1.0 + 0.02 * np.sin(t / 10.0)
    L0 = np.asfortranarray(bg[:, :, None] * illum[None, None, :])
    D = L0 + rng.normal(0.0, 2.0, L0.shape)
    nblk = 6
    y0 = rng.integers(0, max(1, n1 - 10), nblk)
    x0 = rng.integers(0, max(1, n2 - 10), nblk)
    vx = rng.integers(1, 4, nblk)
    for b in range(nblk):
        for tt in range(n3):
            x = int((x0[b] + vx[b] * tt) % max(1, n2 - 10))
            D[y0[b]:y0[b] + 10, x:x + 10, tt] = 250.0
    D = np.asfortranarray(np.clip(D, 0.0, 255.0))
    return (D, L0) if with_truth else D


def make_config(name, with_truth=False, shrink=None):
    """Returns dict(D, r, opts, A0, B0, C0[, L0]) for a BASELINE config.
    `shrink=(n1,n2,n3)` generates the same kind of data at a smaller shape (tests)."""
    n1, n2, n3, r, kind, frac, seed, opts = CONFIGS[name]
    if shrink is not None:
        n1, n2, n3 = shrink
    if kind == "lowrank_sparse":
        out = make_lowrank_sparse(n1, n2, n3, r, frac, seed, with_truth=with_truth)
    elif kind == "traffic":
        out = make_traffic(n1, n2, n3, r, frac, seed, with_truth=with_truth)
    else:
        out = make_video(n1, n2, n3, seed, with_truth=with_truth)
    D, L0 = out if with_truth else (out, None)
    A0, B0, C0 = init_factors(n1, n2, n3, r, 100 + seed)
    res = dict(name=name, D=D, r=r, opts=dict(opts), A0=A0, B0=B0, C0=C0, shape=(n1, n2, n3))
    if with_truth:
        res["L0"] = L0
    return res


def rre(Xhat, X):
    """evaluate() of traffic_triple_comparison.m:194-202 with an all-true mask."""
    return float(np.linalg.norm((Xhat - X).ravel(order="K")) / np.linalg.norm(X.ravel(order="K")))


def slab_bounds(n3, nranks):
    """Contiguous mode-3 slabs [t0,t1) per rank; the first n3 % nranks ranks get one extra
    slice (300 over 8 ranks -> 4 x 38 + 4 x 37).  Mirrors tritd_slab_bounds() in the C ABI."""
    base, extra = divmod(n3, nranks)
    bounds, t0 = [], 0
    for g in range(nranks):
        t1 = t0 + base + (1 if g < extra else 0)
        bounds.append((t0, t1))
        t0 = t1
    return bounds
