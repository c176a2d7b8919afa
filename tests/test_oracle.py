"""CPU tests of the oracle: definitional pins from the reference, algebraic identities the CUDA
design relies on, and the committed golden vectors."""
import os

import numpy as np
import pytest

import make_golden
import tritd_oracle as orc
import tritd_oracle_sharded as orcs
from conftest import rel_err
from tritd import synth


def _rand_factors(n1, n2, n3, r, seed=0):
    rng = np.random.default_rng(seed)
    return (np.asfortranarray(rng.standard_normal((n1, r, r))), np.asfortranarray(rng.standard_normal((r, n2, r))),
            np.asfortranarray(rng.standard_normal((r, r, n3))))


@pytest.mark.parametrize("shape,r", [((7, 6, 5), 3), ((4, 9, 3), 2), ((5, 5, 5), 1)])
def test_definitional_loops(shape, r):
    """Vectorised buildF/G/H/triple_product == the scalar loops the reference holds as comments
    (buildF.m:5-16, buildG.m:5-16, buildH.m:5-16, origin_triple_tensor/triple_decomp_ADMM.m:125-143)."""
    A, B, C = _rand_factors(*shape, r)
    assert np.array_equal(orc.buildF(B, C), orc.buildF_loops(B, C))
    assert np.array_equal(orc.buildG(A, C), orc.buildG_loops(A, C))
    assert np.array_equal(orc.buildH(A, B), orc.buildH_loops(A, B))
    assert rel_err(orc.triple_product(A, B, C), orc.triple_product_loops(A, B, C)) < 1e-14


def test_unfold_index_maps():
    n1, n2, n3 = 4, 3, 5
    X = np.asfortranarray(np.arange(n1 * n2 * n3, dtype=float).reshape((n1, n2, n3), order="F"))
    X2, X3 = orc.unfold(X, 2), orc.unfold(X, 3)
    for i in range(n1):
        for j in range(n2):
            for t in range(n3):
                assert orc.unfold(X, 1)[i, j + t * n2] == X[i, j, t]
                assert X2[j, i + t * n1] == X[i, j, t]
                assert X3[t, i + j * n1] == X[i, j, t]
    with pytest.raises(ValueError):
        orc.unfold(X, 4)


def test_cp_identities():
    """X_(k) M' is an MTTKRP and M M' is a Hadamard product of small Grams."""
    n1, n2, n3, r = 9, 8, 7, 3
    A, B, C = _rand_factors(n1, n2, n3, r, 1)
    T = np.asfortranarray(np.random.default_rng(2).standard_normal((n1, n2, n3)))
    A1, B2, C3 = orc.factors_to_unfolded(A, B, C)
    F, G, H = orc.buildF(B, C), orc.buildG(A, C), orc.buildH(A, B)
    assert rel_err(orc.mttkrp(T, A1, B2, C3, 1), orc.unfold(T, 1) @ F.T) < 1e-13
    assert rel_err(orc.mttkrp(T, A1, B2, C3, 2), orc.unfold(T, 2) @ G.T) < 1e-13
    assert rel_err(orc.mttkrp(T, A1, B2, C3, 3), orc.unfold(T, 3) @ H.T) < 1e-13
    assert rel_err((B2.T @ B2) * (C3.T @ C3), F @ F.T) < 1e-13
    assert rel_err((A1.T @ A1) * (C3.T @ C3), G @ G.T) < 1e-13
    assert rel_err((A1.T @ A1) * (B2.T @ B2), H @ H.T) < 1e-13


def test_soft_threshold_and_pinv():
    x = np.array([-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 3.0])
    assert np.array_equal(orc.soft_threshold(x, 1.0), np.array([-2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.0]))
    G = np.diag([1.0, 1e-20, 2.0])
    P = orc.pinv_matlab(G)
    assert P[1, 1] == 0.0 and P[0, 0] == 1.0 and P[2, 2] == 0.5       # MATLAB cutoff zeroes the tiny one
    assert orc.pinv_truncations(G) == 1


def test_missing_opts_field_is_an_error():
    w = synth.make_config("cfg1", shrink=(6, 5, 4))
    o = dict(w["opts"]); del o["lambda2"]
    with pytest.raises(KeyError, match="lambda2"):
        orc.triple_decomp_ADMM(w["D"], 2, o, *synth.init_factors(6, 5, 4, 2, 0))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_matches_golden(name, golden_dir):
    """The committed vectors pin the oracle (regenerated inputs, stored outputs)."""
    D, r, o, A0, B0, C0 = make_golden.case_inputs(name)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    assert np.allclose([D.sum(), np.abs(D).sum()], g["D_checksum"], rtol=1e-12)
    A, B, C, O, eh = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    assert len(eh) == len(g["errHist"])
    assert rel_err(eh, g["errHist"]) < 1e-10
    for x, key in ((A, "A"), (B, "B"), (C, "C"), (O, "O")):
        assert rel_err(x, g[key]) < 1e-9, key


def test_stop_rule_fires_with_margin(golden_dir):
    g = np.load(os.path.join(golden_dir, "stop_30x30x30_r3.npz"))
    eh = g["errHist"]
    assert 1 < len(eh) < 100
    rel = np.abs(np.diff(eh)) / eh[:-1]
    assert rel[-1] < 0.5 * 2e-2 and np.all(rel[:-1] > 1.05 * 2e-2)      # robust to 1e-12-level noise


def test_sharded_restatement_equals_plain_oracle():
    """The CP / Hadamard / shared-P form the CUDA library uses is the same algorithm."""
    D, r, o, A0, B0, C0 = make_golden.case_inputs("small_40x36x24_r5")
    o["maxIter"] = 8
    A, B, C, O, eh = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    A1, B2, C3, O2, eh2 = orcs.admm_sharded(D, r, o, *orc.factors_to_unfolded(A0, B0, C0))
    assert rel_err(eh2, eh) < 1e-10
    assert rel_err(A1, orc.unfold(A, 1)) < 1e-9 and rel_err(B2, orc.unfold(B, 2)) < 1e-9
    assert rel_err(C3, orc.unfold(C, 3)) < 1e-9 and rel_err(O2, O) < 1e-9


def test_als_runs_and_decreases():
    w = synth.make_config("cfg1", shrink=(12, 11, 10), with_truth=True)
    A, B, C, eh = orc.triple_decomp_ALS(w["L0"], 5, dict(maxIter=15, tol=0.0), w["A0"], w["B0"], w["C0"])
    assert len(eh) == 15 and eh[-1] < eh[0]


def test_synth_slabs_and_bounds():
    full = synth.make_lowrank_sparse(20, 10, 2000, 2, 0.1, 7)
    part = synth.make_lowrank_sparse(20, 10, 2000, 2, 0.1, 7, t0=0, t1=2000)
    assert np.array_equal(full, part)
    assert synth.slab_bounds(300, 8) == [(0, 38), (38, 76), (76, 114), (114, 152), (152, 189), (189, 226), (226, 263), (263, 300)]
    assert synth.slab_bounds(5, 2) == [(0, 3), (3, 5)]


def test_multithreaded_port_equals_oracle():
    """oracle/tritd_oracle_mt.py (the timed CPU baseline of bench.py) makes the same iterates as the numpy oracle."""
    import tritd_oracle_mt as mt
    for cfg, shape, k in (("cfg1", (20, 18, 12), 8), ("cfg3", (24, 32, 10), 6)):
        w = synth.make_config(cfg, shrink=shape)
        o = dict(w["opts"], maxIter=k, tol=0.0)
        ref = orc.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
        res = mt.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"], threads=2)
        for a, b in zip(res, ref):
            assert rel_err(a, b) < 1e-10
    w = synth.make_config("cfg1", shrink=(30, 30, 30))
    o = dict(w["opts"], maxIter=100, tol=2e-2)          # the stopping rule fires at the same iteration
    assert len(mt.triple_decomp_ADMM(w["D"], 3, o, *synth.init_factors(30, 30, 30, 3, 101))[4]) == \
        len(orc.triple_decomp_ADMM(w["D"], 3, o, *synth.init_factors(30, 30, 30, 3, 101))[4])


def test_qi_design_matrices_against_their_scalar_definitions():
    """origin_triple_tensor/buildF|G|H.m restated with the reference's reshape/permute lines = the commented scalar sums;
    A_(1) * F reproduces the six-loop triple product (origin_triple_tensor/triple_product.m)."""
    rng = np.random.default_rng(3)
    n1, n2, n3, r = 5, 4, 3, 2
    A = np.asfortranarray(rng.standard_normal((n1, r, r))); B = np.asfortranarray(rng.standard_normal((r, n2, r)))
    C = np.asfortranarray(rng.standard_normal((r, r, n3)))
    assert np.allclose(orc.buildF_qi(B, C), orc.design_qi_loops(0, B, C), rtol=1e-14, atol=1e-14)
    assert np.allclose(orc.buildG_qi(A, C), orc.design_qi_loops(1, A, C), rtol=1e-14, atol=1e-14)
    assert np.allclose(orc.buildH_qi(A, B), orc.design_qi_loops(2, A, B), rtol=1e-14, atol=1e-14)
    X = np.zeros((n1, n2, n3))
    for i in range(n1):
        for j in range(n2):
            for t in range(n3):
                X[i, j, t] = sum(A[i, q, s] * B[p, j, s] * C[p, q, t] for p in range(r) for q in range(r) for s in range(r))
    assert np.allclose(orc.triple_product_qi(A, B, C), X, rtol=1e-13, atol=1e-13)
    assert np.allclose(np.reshape(np.reshape(A, (n1, r * r), order="F") @ orc.buildF_qi(B, C), (n1, n2, n3), order="F"), X,
                       rtol=1e-13, atol=1e-13)


def test_sparse_pair_is_a_function_of_R3():
    """The identity k_admm's state compression rests on, checked on the ORACLE's own iterates, which keep E and
    Y_O like the reference (triple_decomp_ADMM.m:46-47,:53): with Z = R3 = O + (1/muO)*Y_O_old of an iteration,
    E = Z - clip(Z, +-lambda/muO) bit for bit, and Y_O_new = Y_O_old + muO*(O - E) = muO*clip(Z) up to a few roundings of a
    quantity bounded by lambda."""
    from tritd import synth
    w = synth.make_config("cfg1", shrink=(24, 20, 16))
    opts = dict(w["opts"], maxIter=12, tol=0.0)
    lam, mu0, rho = opts["lambda"], opts["mu"], opts["rho"]
    seen = {"Y_O": np.zeros(w["D"].shape), "n": 0, "dev": 0.0}

    def on_iter(k, A, B, C, O, E, Y_L, Y_O):
        muO = min(mu0 * rho ** (k - 1), mu0 * 1e6)          # the muO iteration k ran with
        thr = lam / muO
        Z = O + (1 / muO) * seen["Y_O"]
        clip = np.minimum(np.maximum(Z, -thr), thr)
        assert np.array_equal(Z - clip, E)                  # soft_threshold(Z, thr), the very same subtraction
        dev = np.abs(muO * clip - Y_O).max()
        assert dev <= 8 * np.finfo(float).eps * lam, dev    # |Y_O| <= lambda
        assert np.abs(Y_O).max() <= lam * (1 + 1e-12)
        seen["Y_O"] = Y_O.copy(); seen["n"] += 1; seen["dev"] = max(seen["dev"], dev)

    orc.triple_decomp_ADMM(w["D"], w["r"], opts, w["A0"], w["B0"], w["C0"], on_iter=on_iter)
    assert seen["n"] == 12


def test_final_rre_fixture_is_the_oracles_outcome():
    """tests/golden/final_rre.json (what bench.py prints next to the GPU run's final RRE) was produced by the multi-threaded
    port; the numpy oracle must give the same outcome of the reference's full cfg1 run: same iteration count, same last
    errHist value and the same RRE = ||triple_product(A,B,C) - L0|| / ||L0|| (traffic_triple_comparison.m:194-199)."""
    import json
    import os
    from tritd import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "final_rre.json")))["cfg1"]
    w = synth.make_config("cfg1", with_truth=True)
    A, B, C, O, eh = orc.triple_decomp_ADMM(w["D"], w["r"], dict(w["opts"], disp=0), w["A0"], w["B0"], w["C0"])
    assert fx["shape"] == list(w["shape"]) and len(eh) == fx["iterations"]
    assert abs(eh[-1] - fx["final_errHist"]) < 1e-6 * fx["final_errHist"]
    rre = synth.rre(orc.triple_product(A, B, C), w["L0"])
    assert abs(rre - fx["RRE"]) < 1e-6 * fx["RRE"] and rre < 1e-7          # the low-rank part is recovered
