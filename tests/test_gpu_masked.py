"""The opt-in extensions of SURVEY 8f rank 2 through the C ABI: the E output and the completion variant
triple_ADMM_masked (named at traffic_triple_comparison.m:53, not shipped by the reference; spec = DESIGN.md 4.6 =
oracle/tritd_oracle.py::triple_ADMM_masked), NaN semantics of the default path, the pinv path inside a full solve."""
import numpy as np
import pytest

import tritd
import tritd_oracle as orc
from conftest import rel_err
from tritd import synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _case(shape, r, frac_missing, seed, iters):
    D, L0 = synth.make_lowrank_sparse(*shape, r, 0.05, seed, with_truth=True)
    F = synth.init_factors(*shape, r, seed + 1)
    m = np.random.default_rng(seed + 2).random(shape) >= frac_missing
    o = dict(synth.TRAFFIC_OPTS, maxIter=iters, tol=0.0)
    return D, L0, F, m, o


def test_E_output_matches_oracle():
    w = synth.make_config("cfg1", shrink=(40, 36, 24))
    o = dict(w["opts"], maxIter=12, tol=0.0)
    A, B, C, O, eh, info = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"], return_info=True, want_E=True)
    ref = orc.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"], return_state=True)
    assert rel_err(info["E"], ref[5]["E"]) < TOL and rel_err(O, ref[3]) < TOL
    # the default 5-output call made the same iterates
    A2, B2, C2, O2, eh2 = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    assert np.array_equal(A, A2) and np.array_equal(O, O2) and np.array_equal(eh, eh2)


@pytest.mark.parametrize("shape,r,frac", [((40, 36, 24), 3, 0.3), ((33, 17, 9), 2, 0.5), ((130, 70, 11), 4, 0.15), ((72, 40, 9), 8, 0.2)])
def test_masked_matches_oracle(shape, r, frac):
    D, L0, F, m, o = _case(shape, r, frac, 21, 15)
    Dg = D.copy(order="F"); Dg[~m] = 1e30                           # values under the mask must not matter
    A, B, C, O, E, out = tritd.triple_ADMM_masked(Dg, m, r, o, *F)
    Ar, Br, Cr, Or, Er, outr = orc.triple_ADMM_masked(D, m, r, o, *F)
    assert len(out["errHist"]) == len(outr["errHist"]) and rel_err(out["errHist"], outr["errHist"]) < TOL
    for x, y in ((A, Ar), (B, Br), (C, Cr), (O, Or), (E, Er)):
        assert rel_err(x, y) < TOL
    assert not O[~m].any() and not E[~m].any()


def test_masked_completes_what_zero_fill_cannot():
    D, L0, F, m, o = _case((40, 36, 24), 3, 0.3, 11, 100)
    A, B, C, O, E, out = tritd.triple_ADMM_masked(D, m, 3, o, *F)
    rre = lambda L, sel: np.linalg.norm((L - L0)[sel]) / np.linalg.norm(L0[sel])   # noqa: E731
    L = tritd.triple_product(A, B, C)
    Az, Bz, Cz, Oz, ehz = tritd.triple_decomp_ADMM(np.where(m, D, 0.0), 3, o, *F)      # the drivers' zero-fill way
    Lz = tritd.triple_product(Az, Bz, Cz)
    assert rre(L, ~m) < 1e-5 and rre(L, ~m) < 0.5 * rre(Lz, ~m)


def test_all_ones_mask_is_the_unmasked_solver_bit_for_bit():
    w = synth.make_config("cfg1", shrink=(48, 40, 12))
    o = dict(w["opts"], maxIter=10, tol=0.0)
    ref = tritd.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    A, B, C, O, E, out = tritd.triple_ADMM_masked(w["D"], np.ones(w["D"].shape, bool), w["r"], o, w["A0"], w["B0"], w["C0"])
    assert np.array_equal(A, ref[0]) and np.array_equal(B, ref[1]) and np.array_equal(C, ref[2])
    assert np.array_equal(O, ref[3]) and np.array_equal(out["errHist"], ref[4])


def test_staged_masked_and_resident_evaluate():
    D, L0, F, m, o = _case((40, 36, 24), 3, 0.3, 11, 40)
    with tritd.Problem(tritd.default_context(), 40, 36, 24, 3) as p:
        p.set_D(D); p.set_mask(m); p.init(o, *F)
        assert p.iterate() == 40
        res = p.get()
        E = p.get_E()
        rmse, nrmse = p.evaluate(L0, ~m)                             # RRE on the held-out entries, factors stay on the device
    Ar, Br, Cr, Or, Er, outr = orc.triple_ADMM_masked(D, m, 3, o, *F)
    assert rel_err(res["A"], Ar) < TOL and rel_err(E, Er) < TOL
    Lr = outr["L"]
    assert abs(nrmse - np.linalg.norm((Lr - L0)[~m]) / np.linalg.norm(L0[~m])) < 1e-8 * max(1.0, nrmse) + 1e-12


def test_nan_in_default_mode_follows_matlab():
    """MATLAB: sign(NaN).*max(..) = NaN and pinv of a NaN matrix is an error ("Input to SVD must not contain NaN or
    Inf") -- no silently finite E; soft_threshold(NaN) = NaN."""
    x = np.array([np.nan, -2.0, 0.5, np.nan, 3.0])
    out = tritd.soft_threshold(x.reshape(5, 1, 1), 1.0).ravel()
    assert np.isnan(out[0]) and np.isnan(out[3]) and np.array_equal(out[[1, 2, 4]], [-1.0, 0.0, 2.0])
    w = synth.make_config("cfg1", shrink=(20, 18, 12))
    D = w["D"].copy(order="F"); D[3, 4, 5] = np.nan
    with pytest.raises(tritd.TritdError) as e:
        tritd.triple_decomp_ADMM(D, w["r"], dict(w["opts"], maxIter=3, tol=0.0), w["A0"], w["B0"], w["C0"])
    assert e.value.code == 4


def test_pinv_path_inside_a_solve():
    """lambda2 = 0 and duplicated columns in B0 and C0: the ridge systems of update_A and update_B are exactly
    singular; the reference's pinv returns the minimum-norm factors, and so does the solver (no error, no garbage)."""
    shape, r = (30, 28, 26), 3
    D = synth.make_lowrank_sparse(*shape, r, 0.05, 31)
    A0, B0, C0 = synth.init_factors(*shape, r, 32)
    B0[1, :, 0] = B0[0, :, 0]; C0[1, 0, :] = C0[0, 0, :]            # columns k = 0 and k = 1 of B2 and C3 coincide
    o = dict(synth.TRAFFIC_OPTS, lambda2=0.0, maxIter=1, tol=0.0)
    with tritd.Problem(tritd.default_context(), *shape, r) as p:
        p.set_D(D); p.init(o, A0, B0, C0)
        assert p.iterate() == 1
        res = p.get()
        fallbacks, truncated = p.pinv_stats()
    ref = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    assert fallbacks >= 2 and truncated >= 2
    assert rel_err(res["A"], ref[0]) < 1e-8 and rel_err(res["B"], ref[1]) < 1e-7
    assert np.all(np.isfinite(res["C"])) and np.all(np.isfinite(res["errHist"]))
