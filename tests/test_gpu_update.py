"""The factor update in isolation (update_A/B/C after the contraction, triple_decomp_ADMM.m:77-78 / :86 / :93):
X = RHS * pinv(S1 .* S2 + alpha*I) through the solver's own update kernel k_upd -- its row reduction, the ridge
inverse, the apply step and the Gram phase X'X -- against numpy, including the truncating pinv path on rank-deficient
systems, sizes above one wave of CTAs, and a no-hang stress loop over the kernel's inter-CTA hand-shakes."""
import numpy as np
import pytest

import tritd
import tritd_oracle as orc
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _system(n, r, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    R = r * r
    U = rng.standard_normal((3 * R + 5, R)) * scale; V = rng.standard_normal((2 * R + 3, R))
    S1, S2 = U.T @ U, V.T @ V
    rhs = np.asfortranarray(rng.standard_normal((n, R)) * 10)
    return rhs, np.asfortranarray(S1), np.asfortranarray(S2)


@pytest.mark.parametrize("r", [1, 2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("n", [1, 7, 300, 1500, 4096])
def test_update_against_numpy(n, r):
    rhs, S1, S2 = _system(n, r, 100 * r + n % 97)
    alpha = 1e-3
    G = S1 * S2 + alpha * np.eye(r * r)
    X, Gi, XtX, info = tritd.factor_update(rhs, S1, S2, alpha)
    cond = np.linalg.cond(G)
    assert info == (0, 0)                                           # well conditioned: the direct inverse
    assert np.abs(Gi @ G - np.eye(r * r)).max() < 1e-14 * cond * r * r
    assert rel_err(X, rhs @ orc.pinv_matlab(G)) < 1e-15 * cond * 10
    assert rel_err(XtX, X.T @ X) < 1e-13


@pytest.mark.parametrize("r,n", [(3, 40), (5, 333), (8, 1100)])
def test_update_rank_deficient_uses_pinv_cutoff(r, n):
    """Duplicated factor columns and no ridge: G is exactly singular; MATLAB's pinv zeroes the singular values below
    max(size)*eps(sigma_max) and returns the minimum-norm solution -- so must the update."""
    rng = np.random.default_rng(r)
    R = r * r
    U = rng.standard_normal((4 * R, R)); V = rng.standard_normal((3 * R, R))
    U[:, 1] = U[:, 0]; V[:, 1] = V[:, 0]                          # columns 0 and 1 duplicated in both factors
    U[:, R - 1] = U[:, 2]; V[:, R - 1] = V[:, 2]                  # ... and another pair
    S1, S2 = np.asfortranarray(U.T @ U), np.asfortranarray(V.T @ V)
    G = S1 * S2
    rhs = np.asfortranarray(rng.standard_normal((n, R)) @ G)      # right-hand sides in the range of G (as X_(k) M' is)
    X, Gi, XtX, info = tritd.factor_update(rhs, S1, S2, 0.0)
    assert info[0] == 1 and info[1] == orc.pinv_truncations(G) == 2
    assert rel_err(Gi, orc.pinv_matlab(G)) < 1e-9
    assert rel_err(X, rhs @ orc.pinv_matlab(G)) < 1e-9
    assert rel_err(XtX, X.T @ X) < 1e-13


def test_update_tiny_ridge_against_large_scale():
    """update_C's fixed 1e-9 ridge (:93) against sigma_max ~ 1e9 (0..255 video data, over-specified r): the ridge
    direction is below pinv's cutoff and gets truncated instead of being inverted at cond 1e18."""
    r, n = 4, 64
    rng = np.random.default_rng(7)
    R = r * r
    U = rng.standard_normal((5 * R, R)) * 3e2; V = rng.standard_normal((5 * R, R)) * 3e2
    U[:, 3] = U[:, 5]; V[:, 3] = V[:, 5]
    S1, S2 = np.asfortranarray(U.T @ U), np.asfortranarray(V.T @ V)
    G = S1 * S2 + 1e-9 * np.eye(R)
    assert orc.pinv_truncations(G) == 1
    rhs = np.asfortranarray(rng.standard_normal((n, R)) @ (S1 * S2))
    X, Gi, XtX, info = tritd.factor_update(rhs, S1, S2, 1e-9)
    assert info == (1, 1)
    assert rel_err(X, rhs @ orc.pinv_matlab(G)) < 1e-8


def test_update_nonfinite_is_an_error():
    rhs, S1, S2 = _system(10, 3, 1)
    S1[2, 2] = np.nan
    with pytest.raises(tritd.TritdError) as e:
        tritd.factor_update(rhs, S1, S2, 1e-3)
    assert e.value.code == 4


def test_update_stress_no_hang():
    """k_upd's hand-shakes (block 0 -> row CTAs -> Gram CTAs) under load: n far above one wave, r = 8, 1000 launches."""
    rhs, S1, S2 = _system(4096, 8, 5)
    ref = None
    for it in range(1000):
        X, Gi, XtX, info = tritd.factor_update(rhs, S1, S2, 1e-2)
        if ref is None:
            ref = (X.copy(), XtX.copy())
            assert rel_err(X, rhs @ np.linalg.inv(S1 * S2 + 1e-2 * np.eye(64))) < 1e-10
        elif it % 100 == 0:
            assert np.array_equal(X, ref[0]) and np.array_equal(XtX, ref[1])        # run-to-run bit-identical
