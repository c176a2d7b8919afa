"""Parity pinned to the reference's own source text.

tests/golden/ref_m/*.npz are outputs of the UNMODIFIED .m files of /root/reference executed by the MATLAB-subset
interpreter oracle/mlab.py (generator: tests/golden/make_ref_golden.py; randn at triple_decomp_ADMM.m:23 shadowed to
inject A0, B0, C0).  Here
  * the numpy oracle is checked against them (CPU; tolerance 1e-10, observed 1e-13: summation-order noise only),
  * the fixtures are re-derived live from /root/reference when it exists (build container) so they cannot drift,
  * the CUDA path is checked against them through the C ABI (-m gpu; north_star tolerance 1e-8),
  * a dump produced by real MATLAB / Octave (tools/reference_dump.m -> tests/golden/matlab_*.mat) is consumed when
    present.
MathWorks' built-ins (pinv, mtimes, norm) are NOT pinned by the interpreter (LAPACK / OpenBLAS through numpy)."""
import glob
import os

import numpy as np
import pytest

import make_golden
import make_ref_golden as mrg
import tritd_oracle as orc
from conftest import rel_err
from tritd import synth

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_m")
HAVE_REF = os.path.isdir(mrg.FR)
TOL_CPU = 1e-10      # oracle vs interpreter: both float64 numpy, differences are summation order in dgemm / norm
TOL_GPU = 1e-8       # north_star: "factor and reconstruction relative error within 1e-8 after a fixed iteration count"


def _load(name):
    return np.load(os.path.join(REF_DIR, name + ".npz"))


def _check_O(O, g, tol):
    s = int(g["O_stride"][0])
    assert rel_err(O[:, :, ::s], g["O"]) < tol
    assert abs(np.linalg.norm(O.ravel()) - g["O_norm"][0]) <= tol * g["O_norm"][0]


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle against the reference source
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(mrg.ADMM_CASES))
def test_oracle_matches_reference_source(name, capsys):
    D, r, o, A0, B0, C0 = mrg.case_inputs(name)
    g = _load(name)
    assert np.allclose([D.sum(), np.abs(D).sum()], g["D_checksum"], rtol=1e-12)
    A, B, C, O, eh = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    printed = capsys.readouterr().out
    assert len(eh) == len(g["errHist"])                       # identical iteration count
    assert rel_err(eh, g["errHist"]) < TOL_CPU
    for x, key in ((A, "A"), (B, "B"), (C, "C")):
        assert rel_err(x, g[key]) < TOL_CPU, key
    _check_O(O, g, TOL_CPU)
    assert printed == str(g["printed"])                       # "Iter %d, errL=%.2e, errO=%.2e" lines, :60-62


@pytest.mark.parametrize("name", sorted(mrg.ALS_CASES))
def test_oracle_als_matches_reference_source(name):
    X, r, o, A0, B0, C0 = make_golden.case_inputs(name)
    g = _load(name)
    A, B, C, eh = orc.triple_decomp_ALS(X, r, o, A0, B0, C0)
    assert len(eh) == len(g["errHist"])
    assert rel_err(eh, g["errHist"]) < TOL_CPU
    for x, key in ((A, "A"), (B, "B"), (C, "C")):
        assert rel_err(x, g[key]) < 1e-8, key                 # ALS ridge 1e-9: cond ~1e6 amplifies dgemm-order noise


def _helper_inputs():
    n1, n2, n3, r = 7, 6, 5, 3
    A, B, C = synth.init_factors(n1, n2, n3, r, 7)
    X = np.asfortranarray(np.random.Generator(np.random.PCG64(8)).standard_normal((n1, n2, n3)))
    gt = np.asfortranarray(np.random.Generator(np.random.PCG64(9)).standard_normal((n1, n2, n3)))
    mask = np.asfortranarray(np.random.Generator(np.random.PCG64(10)).random((n1, n2, n3)) < 0.4)
    return A, B, C, X, gt, mask


def test_oracle_helpers_match_reference_source():
    A, B, C, X, gt, mask = _helper_inputs()
    g = _load("helpers_7x6x5_r3")
    for mode in (1, 2, 3):
        assert np.array_equal(orc.unfold(X, mode), g[f"unfold{mode}"])
    assert np.array_equal(orc.buildF(B, C), g["buildF"])
    assert np.array_equal(orc.buildG(A, C), g["buildG"])
    assert np.array_equal(orc.buildH(A, B), g["buildH"])
    assert rel_err(orc.triple_product(A, B, C), g["triple_product"]) < 1e-14
    assert np.array_equal(orc.soft_threshold(X, 0.7), g["soft_threshold"])
    assert rel_err(orc.buildF_qi(B, C), g["qi_buildF"]) < 1e-14
    assert rel_err(orc.buildG_qi(A, C), g["qi_buildG"]) < 1e-14
    assert rel_err(orc.buildH_qi(A, B), g["qi_buildH"]) < 1e-14
    # the reference's RPAS == Kronecker claim (README.md:43): kronF.m gives the same rows with (q,s) swapped
    K = g["qi_kronF"].reshape((3, 3, -1), order="F").transpose(1, 0, 2).reshape((9, -1), order="F")
    assert rel_err(K, g["qi_buildF"]) < 1e-13
    Xhat = orc.triple_product(A, B, C)
    rm = np.linalg.norm(Xhat[mask] - gt[mask]); nrm = rm / np.linalg.norm(gt[mask])
    assert np.allclose([rm, nrm], g["evaluate_masked"], rtol=1e-13)
    assert np.allclose(synth.rre(Xhat, gt), g["evaluate_all"][1], rtol=1e-13)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference exists only in the build container")
@pytest.mark.parametrize("name", ["tiny_7x6x5_r3", "odd_33x17x9_r2", "stop_30x30x30_r3"])
def test_fixtures_are_live_outputs_of_the_reference_source(name):
    """Re-executes fast_robust_triple_tensor/triple_decomp_ADMM.m from /root/reference and compares with the
    committed fixture; also checks MATLAB's resolution order (local functions shadow the path)."""
    D, r, o, A0, B0, C0 = mrg.case_inputs(name)
    A, B, C, O, eh, text, it = mrg.run_admm(D, r, o, A0, B0, C0)
    g = _load(name)
    assert len(eh) == len(g["errHist"]) and rel_err(eh, g["errHist"]) < 1e-12
    assert rel_err(A, g["A"]) < 1e-11 and rel_err(B, g["B"]) < 1e-11 and rel_err(C, g["C"]) < 1e-11
    res = mrg.resolution_summary(it)
    solver = "fast_robust_triple_tensor/triple_decomp_ADMM.m"
    for fn in ("update_A", "update_B", "update_C", "buildG", "buildH", "reshape_A_from_A1"):
        assert res[fn] == [solver]
    assert res["triple_product"] == ["fast_robust_triple_tensor/triple_product.m"]
    assert res["buildF"] == ["fast_robust_triple_tensor/buildF.m", solver]      # file version only via triple_product.m:6
    assert res["unfold"] == [solver, "fast_robust_triple_tensor/unfold.m"]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference exists only in the build container")
def test_helper_fixtures_are_live_outputs_of_the_reference_source():
    live = mrg.helper_outputs()
    g = _load("helpers_7x6x5_r3")
    for k in live:
        assert np.allclose(live[k], g[k], rtol=1e-13, atol=0), k


# ---------------------------------------------------------------------------------------------------------------
# a dump from real MATLAB / Octave, when somebody has produced one (tools/reference_dump.m)
# ---------------------------------------------------------------------------------------------------------------
MATLAB_DUMPS = sorted(glob.glob(os.path.join(os.path.dirname(REF_DIR), "matlab_*.mat")))


@pytest.mark.skipif(not MATLAB_DUMPS, reason="no tests/golden/matlab_*.mat (run tools/reference_dump.m under MATLAB / Octave)")
@pytest.mark.parametrize("path", MATLAB_DUMPS)
def test_oracle_matches_matlab_dump(path):
    from scipy.io import loadmat
    m = loadmat(path)
    D, r = np.asfortranarray(m["D"]), int(m["r"].ravel()[0])
    o = {k: float(m["opts"][k][0, 0].ravel()[0]) for k in ("mu", "rho", "lambda", "lambda2", "maxIter", "tol")}
    o["maxIter"] = int(o["maxIter"]); o["disp"] = 0
    A, B, C, O, eh = orc.triple_decomp_ADMM(D, r, o, m["A0"], m["B0"], m["C0"])
    assert len(eh) == m["errHist"].size
    assert rel_err(eh, m["errHist"].ravel()) < 1e-8
    for x, key in ((A, "A"), (B, "B"), (C, "C"), (O, "O")):
        assert rel_err(x, m[key].reshape(x.shape, order="F")) < 1e-8, key


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path (through the C ABI) against the reference source
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mrg.ADMM_CASES))
def test_gpu_matches_reference_source(name, capfd):
    import tritd
    D, r, o, A0, B0, C0 = mrg.case_inputs(name)
    g = _load(name)
    A, B, C, O, eh = tritd.triple_decomp_ADMM(D, r, o, A0, B0, C0)
    printed = capfd.readouterr().out
    assert len(eh) == len(g["errHist"])                       # identical iteration count (stop case: the rule fires)
    assert rel_err(eh, g["errHist"]) < TOL_GPU
    for x, key in ((A, "A"), (B, "B"), (C, "C")):
        assert rel_err(x, g[key]) < TOL_GPU, key
    _check_O(O, g, TOL_GPU)
    assert printed == str(g["printed"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mrg.ALS_CASES))
def test_gpu_als_matches_reference_source(name):
    import tritd
    X, r, o, A0, B0, C0 = make_golden.case_inputs(name)
    g = _load(name)
    A, B, C, eh = tritd.triple_decomp_ALS(X, r, dict(o, disp=0), A0, B0, C0)
    assert len(eh) == len(g["errHist"])
    assert rel_err(eh, g["errHist"]) < TOL_GPU
    L, Lr = orc.triple_product(A, B, C), orc.triple_product(g["A"], g["B"], g["C"])
    assert rel_err(L, Lr) < 1e-7                              # ridge 1e-9 (:27,:32,:37): compare the reconstruction


@pytest.mark.gpu
def test_gpu_helpers_match_reference_source():
    import tritd
    A, B, C, X, gt, mask = _helper_inputs()
    g = _load("helpers_7x6x5_r3")
    for mode in (1, 2, 3):
        assert np.array_equal(tritd.unfold(X, mode), g[f"unfold{mode}"])
    assert np.array_equal(tritd.buildF(B, C), g["buildF"])
    assert np.array_equal(tritd.buildG(A, C), g["buildG"])
    assert np.array_equal(tritd.buildH(A, B), g["buildH"])
    assert np.array_equal(tritd.soft_threshold(X, 0.7), g["soft_threshold"])
    assert rel_err(tritd.triple_product(A, B, C), g["triple_product"]) < 1e-13
    assert rel_err(tritd.buildF_qi(B, C), g["qi_buildF"]) < 1e-13
    assert rel_err(tritd.buildG_qi(A, C), g["qi_buildG"]) < 1e-13
    assert rel_err(tritd.buildH_qi(A, B), g["qi_buildH"]) < 1e-13
    rm, nrm = tritd.evaluate(A, B, C, gt, mask)
    assert np.allclose([rm, nrm], g["evaluate_masked"], rtol=1e-12)
    rm, nrm = tritd.evaluate(A, B, C, gt)
    assert np.allclose([rm, nrm], g["evaluate_all"], rtol=1e-12)
