"""The MEX gateways against the REAL library on the B200: gateway source + the mock MEX runtime of tests/mex_mock (see
tests/test_mex_gateway.py, which executes the same gateways on the CPU against a test double) linked with libtritd.so, one
MATLAB-style call each, results checked against the oracle.  The file sorts last on purpose: the gateways point the
library's print sink at mexPrintf (restored at the end of each test)."""
import numpy as np
import pytest

import tritd
import tritd_oracle as orc
from conftest import rel_err
from test_mex_gateway import Gateway, build_gateway
from tritd import synth

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _flat(*arrays):
    return np.concatenate([np.asarray(a, dtype=np.float64).ravel(order="F") for a in arrays])


def admm_case(g, nlhs=7):
    """[A,B,C,O,errHist,L,E] = triple_decomp_ADMM(D, r, opts) on the workload smoke() uses; randn hands out its A0, B0, C0"""
    w = synth.make_config("cfg1", shrink=(40, 36, 24))
    g.randn_source(_flat(w["A0"], w["B0"], w["C0"]))
    opts = dict(w["opts"], maxIter=20, tol=0.0, disp=True)
    out, err = g.call(nlhs, g.double(w["D"]), g.double(w["r"]), g.struct(opts))
    return w, opts, out, err


def test_admm_gateway_on_the_gpu(tmp_path):
    g = Gateway(build_gateway(tmp_path, "triple_decomp_ADMM", real=True), tmp_path)
    try:
        w, opts, out, err = admm_case(g)
        assert err is None, err
        n1, n2, n3 = w["D"].shape
        r = w["r"]
        assert g.randn_calls() == [(n1, r, r), (r, n2, r), (r, r, n3)]
        A, B, C, O, eh, L, E = out
        ref = orc.triple_decomp_ADMM(w["D"], r, dict(opts, disp=0), w["A0"], w["B0"], w["C0"], return_state=True)
        Ar, Br, Cr, Or, ehr, st = ref[0], ref[1], ref[2], ref[3], ref[4], ref[5]
        assert eh.shape == (20, 1) and rel_err(eh[:, 0], ehr) < TOL
        for x, y in ((A, Ar), (B, Br), (C, Cr), (O, Or), (L, orc.triple_product(Ar, Br, Cr)), (E, st["E"])):
            assert x.shape == y.shape and rel_err(x, y) < TOL
        lines = g.printed().splitlines()                       # opts.disp: the reference's progress line, through mexPrintf
        assert len(lines) == 2 and lines[0].startswith("Iter 10, errL=") and lines[1].startswith("Iter 20, errL=")
        # a second call reuses the locked context (and the cached device state): same results
        w2, _, out2, err2 = admm_case(g, nlhs=5)
        assert err2 is None and g.lib.mock_locks() == 1
        assert np.array_equal(out2[0], A) and np.array_equal(out2[4], eh)
        assert g.lib.mock_run_atexit() == 1                    # the at-exit hook destroys the context
    finally:
        tritd.set_print(None)


def test_triple_product_gateway_on_the_gpu(tmp_path):
    g = Gateway(build_gateway(tmp_path, "triple_product", real=True), tmp_path)
    rng = np.random.default_rng(2)
    n1, n2, n3, r = 33, 17, 9, 2
    A, B, C = rng.standard_normal((n1, r, r)), rng.standard_normal((r, n2, r)), rng.standard_normal((r, r, n3))
    out, err = g.call(1, g.double(A), g.double(B), g.double(C))
    assert err is None, err
    assert out[0].shape == (n1, n2, n3) and rel_err(out[0], orc.triple_product(A, B, C)) < 1e-10
    assert g.lib.mock_run_atexit() == 1
