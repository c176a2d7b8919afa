"""GPU parity of triple_decomp_ALS (reference: fast_robust_triple_tensor/triple_decomp_ALS.m)
through the C ABI against the oracle restatement and the committed golden vectors."""
import os

import numpy as np
import pytest

import make_golden
import tritd
import tritd_oracle as orc
from conftest import rel_err
from tritd import synth

pytestmark = pytest.mark.gpu
TOL = 1e-8      # north_star tolerance for factors / reconstruction / error history


def _cmp(res, ref):
    A, B, C, eh = res
    Ar, Br, Cr, ehr = ref
    assert len(eh) == len(ehr), (len(eh), len(ehr))
    errs = dict(errHist=rel_err(eh, ehr), L=rel_err(orc.triple_product(A, B, C), orc.triple_product(Ar, Br, Cr)),
                A=rel_err(A, Ar), B=rel_err(B, Br), C=rel_err(C, Cr))
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("name", sorted(make_golden.ALS_CASES))
def test_als_golden(name, golden_dir):
    X, r, o, A0, B0, C0 = make_golden.case_inputs(name)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    A, B, C, eh = tritd.triple_decomp_ALS(X, r, o, A0, B0, C0, disp=False)
    _cmp((A, B, C, eh), (g["A"], g["B"], g["C"], g["errHist"]))


@pytest.mark.parametrize("shape,r,k", [((40, 36, 24), 5, 6), ((33, 17, 9), 2, 10), ((64, 48, 20), 8, 4), ((130, 70, 11), 4, 5)])
def test_als_trajectory(shape, r, k):
    D = synth.make_lowrank_sparse(*shape, r, 0.05, 11)
    A0, B0, C0 = synth.init_factors(*shape, r, 12)
    o = dict(maxIter=k, tol=0.0)
    _cmp(tritd.triple_decomp_ALS(D, r, o, A0, B0, C0, disp=False), orc.triple_decomp_ALS(D, r, o, A0, B0, C0))


def test_als_stops_like_the_reference():
    """The relative-change rule fires before the updates of that iteration: same iteration count, same factors."""
    shape, r = (36, 30, 28), 3
    D = synth.make_lowrank_sparse(*shape, r, 0.05, 21)
    A0, B0, C0 = synth.init_factors(*shape, r, 22)
    o = dict(maxIter=80, tol=2e-3)
    ref = orc.triple_decomp_ALS(D, r, o, A0, B0, C0)
    assert 2 < len(ref[3]) < 80
    res = tritd.triple_decomp_ALS(D, r, o, A0, B0, C0, disp=False)
    _cmp(res, ref)


def test_als_recovers_exact_low_rank():
    shape, r = (30, 26, 22), 2
    D = synth.make_lowrank_sparse(*shape, r, 0.0, 31)
    A0, B0, C0 = synth.init_factors(*shape, r, 32)
    A, B, C, eh = tritd.triple_decomp_ALS(D, r, dict(maxIter=60, tol=1e-12), A0, B0, C0, disp=False)
    assert eh[-1] < eh[0] and np.all(np.diff(eh) < 1e-9)          # ALS never increases the fit error


def test_als_missing_field_and_progress_line(capfd):
    D = synth.make_lowrank_sparse(12, 10, 8, 2, 0.0, 41)
    with pytest.raises(KeyError, match="Unrecognized field name"):
        tritd.triple_decomp_ALS(D, 2, dict(maxIter=3))
    tritd.triple_decomp_ALS(D, 2, dict(maxIter=10, tol=0.0), rng=np.random.default_rng(1))
    out = capfd.readouterr().out
    assert "Iteration 5, relative error = " in out and "Iteration 10, relative error = " in out
