"""GPU parity, kernel by kernel, through the C ABI against the numpy oracle (SURVEY 4 tiers i-ii)."""
import numpy as np
import pytest

import tritd
import tritd_oracle as orc
from conftest import rel_err
from tritd import synth

pytestmark = pytest.mark.gpu

# float64 contractions in a different summation order: relative error ~ sqrt(K)*eps
TOL = 1e-12


def _factors(n1, n2, n3, r, seed):
    rng = np.random.default_rng(seed)
    return (np.asfortranarray(rng.standard_normal((n1, r, r))), np.asfortranarray(rng.standard_normal((r, n2, r))),
            np.asfortranarray(rng.standard_normal((r, r, n3))))


SHAPES = [((7, 6, 5), 3), ((33, 17, 9), 2), ((16, 32, 4), 1), ((40, 36, 24), 5), ((130, 70, 11), 4),
          ((129, 33, 6), 6), ((64, 48, 10), 7), ((50, 50, 50), 5), ((72, 40, 9), 8), ((257, 65, 5), 5)]


@pytest.mark.parametrize("shape,r", SHAPES)
def test_triple_product(shape, r):
    A, B, C = _factors(*shape, r, 1)
    assert rel_err(tritd.triple_product(A, B, C), orc.triple_product(A, B, C)) < TOL


def test_triple_product_definitional():
    """Against the five-nested-loop definition (origin_triple_tensor/triple_decomp_ADMM.m:125-143)."""
    A, B, C = _factors(7, 6, 5, 3, 2)
    assert rel_err(tritd.triple_product(A, B, C), orc.triple_product_loops(A, B, C)) < TOL


@pytest.mark.parametrize("shape,r", SHAPES)
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_mttkrp(shape, r, mode):
    """X_(k) * M' of update_A/B/C (triple_decomp_ADMM.m:78,:86,:93) with M materialised by the oracle."""
    A, B, C = _factors(*shape, r, 3)
    X = np.asfortranarray(np.random.default_rng(4).standard_normal(shape))
    M = {1: orc.buildF(B, C), 2: orc.buildG(A, C), 3: orc.buildH(A, B)}[mode]
    ref = orc.unfold(X, mode) @ M.T
    assert rel_err(tritd.mttkrp(X, A, B, C, mode), ref) < TOL


@pytest.mark.parametrize("shape", [(7, 6, 5), (33, 17, 9), (64, 64, 3), (100, 31, 40)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_unfold_bit_exact(shape, mode):
    X = np.asfortranarray(np.random.default_rng(5).standard_normal(shape))
    assert np.array_equal(tritd.unfold(X, mode), orc.unfold(X, mode))


@pytest.mark.parametrize("shape,r", [((7, 6, 5), 3), ((33, 17, 9), 2), ((20, 30, 10), 5), ((9, 8, 7), 8)])
def test_build_FGH_bit_exact(shape, r):
    """One multiply per entry, so the result is bit-identical to the reference formula
    (and to the commented scalar loops buildF/G/H.m:5-16)."""
    A, B, C = _factors(*shape, r, 6)
    assert np.array_equal(tritd.buildF(B, C), orc.buildF(B, C))
    assert np.array_equal(tritd.buildG(A, C), orc.buildG(A, C))
    assert np.array_equal(tritd.buildH(A, B), orc.buildH(A, B))
    if max(shape) < 10:
        assert np.array_equal(tritd.buildF(B, C), orc.buildF_loops(B, C))


def test_soft_threshold_bit_exact():
    x = np.random.default_rng(7).standard_normal(100003) * 3
    x[:5] = [0.0, 1.0, -1.0, 1.0 + 1e-16, -0.0]
    for lam in (0.0, 1.0, 2.5):
        assert np.array_equal(tritd.soft_threshold(x, lam), orc.soft_threshold(x, lam))


def test_unsupported_rank_is_an_error():
    A, B, C = _factors(4, 4, 4, 9, 8)
    with pytest.raises(tritd.TritdError) as ei:
        tritd.triple_product(A, B, C)
    assert ei.value.code == 5


@pytest.mark.parametrize("shape,r", [((37, 29, 11), 3), ((64, 50, 20), 5)])
def test_evaluate_on_device(shape, r):
    """evaluate() of the drivers (traffic_triple_comparison.m:194-202) on the device = numpy on the oracle's reconstruction."""
    rng = np.random.default_rng(5)
    A0, B0, C0 = synth.init_factors(*shape, r, 6)
    gt = np.asfortranarray(rng.standard_normal(shape))
    Xhat = orc.triple_product(A0, B0, C0)
    rmse, nrmse = tritd.evaluate(A0, B0, C0, gt)
    ref = np.linalg.norm((Xhat - gt).ravel())
    assert abs(rmse - ref) < 1e-11 * ref and abs(nrmse - ref / np.linalg.norm(gt.ravel())) < 1e-11 * nrmse
    mask = rng.random(shape) < 0.3
    rmse, nrmse = tritd.evaluate(A0, B0, C0, gt, mask)
    ref = np.linalg.norm(Xhat[mask] - gt[mask])
    assert abs(rmse - ref) < 1e-11 * ref and abs(nrmse - ref / np.linalg.norm(gt[mask])) < 1e-11 * nrmse


@pytest.mark.parametrize("shape,r", [((7, 6, 5), 3), ((33, 17, 9), 2), ((40, 36, 24), 5), ((20, 12, 9), 8)])
def test_qi_design_matrices_and_product(shape, r):
    """Design matrices / product of the original (Qi) model (origin_triple_tensor/buildF|G|H.m, triple_product.m)."""
    A, B, C = _factors(*shape, r, 9)
    assert rel_err(tritd.buildF_qi(B, C), orc.buildF_qi(B, C)) < TOL
    assert rel_err(tritd.buildG_qi(A, C), orc.buildG_qi(A, C)) < TOL
    assert rel_err(tritd.buildH_qi(A, B), orc.buildH_qi(A, B)) < TOL
    assert rel_err(tritd.triple_product_qi(A, B, C), orc.triple_product_qi(A, B, C)) < TOL
