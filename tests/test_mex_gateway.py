"""The MEX gateways EXECUTED (not merely syntax-checked): each gateway source under <package>/mex is compiled together with
tests/mex_mock/mexmock.c -- a mock of the part of MATLAB's MEX runtime the gateways use -- into a shared object that these
tests drive through ctypes: argument validation and error identifiers (reject complex / sparse / single / non-3-D,
n3 = 1 arrives as 2-D, missing opts field -> MATLAB's 'Unrecognized field name', unknown fields ignored, nlhs <= 5 honoured),
the order and sizes of the randn calls (fast_robust_triple_tensor/triple_decomp_ADMM.m:23, triple_decomp_ALS.m:8-10), the shapes of
the outputs, errHist trimmed to the executed iterations (:68), the progress line through mexPrintf (:60-62), context caching
under mexLock / mexAtExit.

Without a GPU the solve itself is played by tests/mex_mock/fake_tritd.c, a test double of the five entry points the gateways
call (recognisable outputs, recorded arguments); one test links the REAL libtritd and checks that the gateway reports the
missing device as tritd:cuda (no CPU fallback).  tests/test_zz_mex_gateway_gpu.py runs the same gateway + mock against the
real library on the B200."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import tritd_oracle as orc
from conftest import PKG, ROOT

MOCK = os.path.join(ROOT, "tests", "mex_mock")
GATEWAYS = ("triple_decomp_ADMM", "triple_ADMM_masked", "triple_decomp_ALS", "triple_product")
OPTS = dict(mu=1e-3, rho=1.25, **{"lambda": 1.8}, lambda2=1e-3, maxIter=100, tol=1e-5, disp=0)


def build_gateway(out_dir, name, real):
    """gateway + mock (+ the test double, or the real library) -> <out_dir>/<name>_{fake,real}.so"""
    out = os.path.join(str(out_dir), "%s_%s.so" % (name, "real" if real else "fake"))
    cmd = ["gcc", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-O1",
           "-I" + os.path.join(PKG, "mex", "stub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(PKG, "mex", name + ".c"), os.path.join(MOCK, "mexmock.c")]
    if real:
        libdir = os.path.join(PKG, "tritd")
        cmd += ["-L" + libdir, "-ltritd", "-Wl,-rpath," + libdir]
    else:
        # -Bsymbolic: the gateway's tritd_* calls bind to the test double inside this object even when the real libtritd
        # has already been loaded into the process with global visibility (the other tests do that)
        cmd += [os.path.join(MOCK, "fake_tritd.c"), "-Wl,-Bsymbolic"]
    subprocess.check_call(cmd + ["-o", out])
    return out


class Gateway:
    """One loaded copy of a gateway .so (a fresh copy per instance, so the gateway's cached context starts empty)."""
    _n = 0

    def __init__(self, so_path, tmp_dir):
        Gateway._n += 1
        mine = os.path.join(str(tmp_dir), "load%d_%s" % (Gateway._n, os.path.basename(so_path)))
        shutil.copy(so_path, mine)
        self.lib = lib = ctypes.CDLL(mine)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        lib.mock_double.restype = vp; lib.mock_double.argtypes = [ctypes.c_int, ctypes.POINTER(sz), vp]
        lib.mock_logical.restype = vp; lib.mock_logical.argtypes = [ctypes.c_int, ctypes.POINTER(sz), vp]
        lib.mock_single_scalar.restype = vp; lib.mock_single_scalar.argtypes = [ctypes.c_float]
        lib.mock_struct.restype = vp
        lib.mxSetField.argtypes = [vp, sz, ctypes.c_char_p, vp]
        lib.mxGetField.restype = vp; lib.mxGetField.argtypes = [vp, sz, ctypes.c_char_p]
        lib.mock_set_flags.argtypes = [vp, ctypes.c_int, ctypes.c_int]
        lib.mock_set_randn_source.argtypes = [vp, sz]
        lib.mock_call.argtypes = [ctypes.c_int, ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(vp)]
        lib.mock_err_id.restype = ctypes.c_char_p; lib.mock_err_msg.restype = ctypes.c_char_p
        lib.mock_printed.restype = ctypes.c_char_p
        lib.mock_randn_dims.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        lib.mock_ndim.argtypes = [vp]; lib.mock_dim.restype = sz; lib.mock_dim.argtypes = [vp, ctypes.c_int]
        lib.mxGetPr.restype = ctypes.POINTER(ctypes.c_double); lib.mxGetPr.argtypes = [vp]
        lib.mxIsStruct.argtypes = [vp]
        self._keep = []

    # ---- building mxArrays -------------------------------------------------------------------------------------
    def double(self, x):
        x = np.asfortranarray(np.asarray(x, dtype=np.float64))
        if x.ndim < 2:
            x = x.reshape((1, -1), order="F") if x.ndim == 1 else x.reshape((1, 1))
        dims = (ctypes.c_size_t * x.ndim)(*x.shape)
        self._keep.append(x)
        return _Handle(self.lib.mock_double(x.ndim, dims, x.ctypes.data_as(ctypes.c_void_p)))

    def logical(self, x):
        x = np.asfortranarray(np.asarray(x).astype(np.uint8))
        if x.ndim < 2:
            x = x.reshape((1, -1), order="F") if x.ndim == 1 else x.reshape((1, 1))
        dims = (ctypes.c_size_t * x.ndim)(*x.shape)
        self._keep.append(x)
        return _Handle(self.lib.mock_logical(x.ndim, dims, x.ctypes.data_as(ctypes.c_void_p)))

    def struct(self, fields):
        """dict -> 1x1 struct; values: an mxArray handle (kept), a bool (logical scalar), anything else (double array)"""
        s = self.lib.mock_struct()
        for k, v in fields.items():
            if isinstance(v, _Handle):
                h = v
            elif isinstance(v, (bool, np.bool_)):
                h = self.logical(np.array([[v]]))
            else:
                h = self.double(v)
            self.lib.mxSetField(s, 0, k.encode(), h.h)
        return _Handle(s)

    def single_scalar(self, v):
        return _Handle(self.lib.mock_single_scalar(v))

    def set_field(self, s, name, h):
        self.lib.mxSetField(s.h, 0, name.encode(), h.h)

    def set_flags(self, h, is_complex, is_sparse):
        self.lib.mock_set_flags(h.h, is_complex, is_sparse)

    def randn_source(self, n, seed=0):
        """what the mocked randn hands out, in call order: n seeded numbers, or the given array"""
        src = np.random.default_rng(seed).standard_normal(n) if np.isscalar(n) else np.ascontiguousarray(n, dtype=np.float64)
        self._keep.append(src)
        self.lib.mock_set_randn_source(src.ctypes.data_as(ctypes.c_void_p), src.size)
        return src

    # ---- calling mexFunction -----------------------------------------------------------------------------------
    def call(self, nlhs, *prhs):
        self.lib.mock_reset()
        nout = max(nlhs, 1)
        plhs = (ctypes.c_void_p * 8)(*([None] * 8))
        args = (ctypes.c_void_p * len(prhs))(*[a.h for a in prhs])
        rc = self.lib.mock_call(nlhs, plhs, len(prhs), args)
        if rc:
            return None, (self.lib.mock_err_id().decode(), self.lib.mock_err_msg().decode())
        return [self.array(plhs[i]) if plhs[i] else None for i in range(nout)], None

    def array(self, h):
        if self.lib.mxIsStruct(h):
            return _Handle(h)
        shape = tuple(self.lib.mock_dim(h, i) for i in range(self.lib.mock_ndim(h)))
        n = int(np.prod(shape))
        buf = np.ctypeslib.as_array(self.lib.mxGetPr(h), shape=(n,)).copy() if n else np.zeros(0)
        return buf.reshape(shape, order="F")

    def field(self, h, name):
        return self.array(self.lib.mxGetField(h.h, 0, name.encode()))

    def randn_calls(self):
        out = []
        for i in range(self.lib.mock_randn_calls()):
            d = (ctypes.c_double * 3)()
            self.lib.mock_randn_dims(i, d)
            out.append(tuple(int(v) for v in d))
        return out

    def printed(self):
        return self.lib.mock_printed().decode()

    def fake(self, name, ctype=ctypes.c_int):
        return ctype.in_dll(self.lib, name)


class _Handle:
    def __init__(self, h):
        self.h = h


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    d = tmp_path_factory.mktemp("mexgw")
    return {g: build_gateway(d, g, real=False) for g in GATEWAYS}, d


def _admm(built):
    sos, d = built
    return Gateway(sos["triple_decomp_ADMM"], d)


def _problem(n1=6, n2=5, n3=4, seed=1):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((n1, n2, n3)))


# ---------------------------------------------------------------------------------------------------------------------
def test_admm_gateway_signature_randn_order_and_outputs(built):
    g = _admm(built)
    n1, n2, n3, r = 6, 5, 4, 2
    D = _problem(n1, n2, n3)
    src = g.randn_source((n1 + n2 + n3) * r * r)
    opts = g.struct(dict(OPTS, alphaA=1.0, alphaB=2.0, origin=3.0))                # unknown fields are ignored (callers pass them)
    out, err = g.call(5, g.double(D), g.double(r), opts)
    assert err is None
    # randn(n1,r,r), randn(r,n2,r), randn(r,r,n3) in this order (:23): rng(0) in the caller gives the reference's factors
    assert g.randn_calls() == [(n1, r, r), (r, n2, r), (r, r, n3)]
    A0 = src[: n1 * r * r].reshape((n1, r, r), order="F")
    B0 = src[n1 * r * r: (n1 + n2) * r * r].reshape((r, n2, r), order="F")
    C0 = src[(n1 + n2) * r * r:].reshape((r, r, n3), order="F")
    A, B, C, O, eh = out
    assert A.shape == (n1, r, r) and B.shape == (r, n2, r) and C.shape == (r, r, n3) and O.shape == (n1, n2, n3)
    assert np.array_equal(A, 2 * A0) and np.array_equal(B, 3 * B0) and np.array_equal(C, 4 * C0)   # (the test double's outputs)
    assert np.array_equal(O, D + 1)
    assert eh.shape == (7, 1) and np.array_equal(eh[:, 0], 1.0 / np.arange(1, 8))                 # errHist(1:k), a column (:68)
    o = g.fake("fake_opts", _Opts)
    assert (o.mu, o.rho, o.lambda_, o.lambda2, o.tol, o.maxIter, o.disp) == (1e-3, 1.25, 1.8, 1e-3, 1e-5, 100, 0)
    assert list(g.fake("fake_shape", ctypes.c_longlong * 4)) == [n1, n2, n3, r]
    assert g.fake("fake_mask_seen").value == 0 and g.fake("fake_have_L").value == 0 and g.fake("fake_have_E").value == 0
    # the context is created once, locked in memory, and released by the at-exit hook
    assert g.fake("fake_created").value == 1 and g.lib.mock_locks() == 1
    out2, err2 = g.call(5, g.double(D), g.double(r), opts)
    assert err2 is None and g.fake("fake_created").value == 1 and g.lib.mock_locks() == 1
    assert g.lib.mock_run_atexit() == 1 and g.fake("fake_destroyed").value == 1


class _Opts(ctypes.Structure):
    _fields_ = [("mu", ctypes.c_double), ("rho", ctypes.c_double), ("lambda_", ctypes.c_double), ("lambda2", ctypes.c_double),
                ("tol", ctypes.c_double), ("maxIter", ctypes.c_int32), ("disp", ctypes.c_int32)]


@pytest.mark.parametrize("nlhs", [0, 1, 2, 3, 4, 5, 6, 7])
def test_admm_gateway_honours_nlhs(built, nlhs):
    g = _admm(built)
    D = _problem()
    g.randn_source(200)
    out, err = g.call(nlhs, g.double(D), g.double(2), g.struct(OPTS))
    assert err is None
    assert out[0] is not None and out[0].shape == (6, 2, 2)                       # nlhs == 0 still returns ans = A
    assert len([x for x in out if x is not None]) == max(nlhs, 1)
    assert g.fake("fake_have_O").value == (nlhs >= 4)                             # O is not even copied back unless asked for
    assert g.fake("fake_have_L").value == (nlhs >= 6) and g.fake("fake_have_E").value == (nlhs >= 7)
    if nlhs >= 6:
        assert np.array_equal(out[5], D + 3)                                      # 6th output: L = triple_product(A,B,C)
    if nlhs >= 7:
        assert np.array_equal(out[6], D + 2)                                      # 7th output: E ("O,E", :12)


def test_admm_gateway_argument_errors(built):
    g = _admm(built)
    D, r, opts = g.double(_problem()), g.double(2), g.struct(OPTS)
    g.randn_source(10000)

    def err_of(nlhs, *a):
        out, err = g.call(nlhs, *a)
        assert out is None
        return err

    assert err_of(5, D, r)[0] == "tritd:nargin"
    assert err_of(8, D, r, opts)[0] == "MATLAB:TooManyOutputs"
    # the reference's own failure for a missing field (opts.rho read at :16-20)
    o = dict(OPTS); del o["rho"]
    assert err_of(5, D, r, g.struct(o)) == ("MATLAB:nonExistentField", 'Unrecognized field name "rho".')
    for bad in (np.zeros((1, 2)), np.zeros((0, 0))):                              # non-scalar / empty double field
        assert err_of(5, D, r, g.struct(dict(OPTS, mu=bad)))[0] == "tritd:opts"
    cplx = g.double(1.0); g.set_flags(cplx, 1, 0)
    s = g.struct(OPTS); g.set_field(s, "tol", cplx)
    assert err_of(5, D, r, s)[0] == "tritd:opts"
    s = g.struct(OPTS); g.set_field(s, "mu", g.single_scalar(0.5))
    assert err_of(5, D, r, s)[0] == "tritd:opts"
    assert err_of(5, D, r, g.double(3.0))[0] == "tritd:opts"                       # opts is not a struct
    # D: full real double, 2-D or 3-D, not empty
    for flags in ((1, 0), (0, 1)):
        Dx = g.double(_problem()); g.set_flags(Dx, *flags)
        assert err_of(5, Dx, r, opts)[0] == "tritd:D"
    assert err_of(5, g.logical(np.ones((3, 3, 3))), r, opts)[0] == "tritd:D"
    assert err_of(5, g.double(np.zeros((3, 3, 3, 2))), r, opts)[0] == "tritd:D"
    assert err_of(5, g.double(np.zeros((3, 0, 3))), r, opts)[0] == "tritd:D"
    # r: a positive integer scalar
    for bad in (2.5, 0.0, -1.0, np.array([[2.0, 2.0]])):
        assert err_of(5, D, g.double(bad), opts)[0] == "tritd:r"
    assert g.fake("fake_created").value == 0                                     # nothing reached the library


def test_admm_gateway_accepts_2d_D_as_n3_equal_1(built):
    g = _admm(built)
    D = np.asfortranarray(np.random.default_rng(3).standard_normal((6, 5)))
    g.randn_source(200)
    out, err = g.call(5, g.double(D), g.double(2), g.struct(OPTS))
    assert err is None
    assert list(g.fake("fake_shape", ctypes.c_longlong * 4)) == [6, 5, 1, 2]
    assert g.randn_calls() == [(6, 2, 2), (2, 5, 2), (2, 2, 1)]
    assert out[2].shape == (2, 2) and out[3].shape == (6, 5)                      # MATLAB drops the trailing singleton


def test_admm_gateway_injected_factors_logical_disp_and_progress_lines(built):
    g = _admm(built)
    n1, n2, n3, r = 6, 5, 4, 2
    rng = np.random.default_rng(5)
    A0, B0, C0 = rng.standard_normal((n1, r, r)), rng.standard_normal((r, n2, r)), rng.standard_normal((r, r, n3))
    g.randn_source(0)
    g.fake("fake_iters").value = 25
    opts = dict(OPTS, A0=A0, B0=B0, C0=C0, disp=True, maxIter=40)
    out, err = g.call(5, g.double(_problem()), g.double(r), g.struct(opts))
    assert err is None and g.randn_calls() == []                                  # no random numbers drawn
    assert np.array_equal(out[0], 2 * np.asfortranarray(A0)) and np.array_equal(out[2], 4 * C0)
    assert out[4].shape == (25, 1)
    assert g.fake("fake_opts", _Opts).disp == 1
    lines = g.printed().splitlines()                                              # the progress line reaches mexPrintf (:60-62)
    assert len(lines) == 2 and lines[0].startswith("Iter 10, errL=") and lines[1].startswith("Iter 20, errL=")
    _, err = g.call(5, g.double(_problem()), g.double(r), g.struct(dict(opts, B0=np.zeros((r, n2 + 1, r)))))
    assert err == ("tritd:opts", "opts.B0 has the wrong size or class.")


@pytest.mark.parametrize("field,value,want", [("ngpu", 4.0, [0, 1, 2, 3]), ("devices", np.array([[2.0, 3.0]]), [2, 3]),
                                              ("device", 5.0, [5]), (None, None, [0])])
def test_admm_gateway_device_selection(built, field, value, want):
    g = _admm(built)
    g.randn_source(200)
    o = dict(OPTS)
    if field:
        o[field] = value
    out, err = g.call(5, g.double(_problem()), g.double(2), g.struct(o))
    assert err is None
    assert g.fake("fake_ndev").value == len(want) and list(g.fake("fake_devs", ctypes.c_int * 8))[: len(want)] == want


def test_admm_gateway_recreates_its_context_when_the_devices_change(built):
    """INTEGRATION.md's own sequence: a single-GPU call, then `opts.ngpu = 8` -- the cached context must not silently stay on
    one GPU; an unchanged request reuses the context, and the lock is taken once."""
    g = _admm(built)
    D, r = g.double(_problem()), g.double(2)
    g.randn_source(10000)
    devs = lambda: list(g.fake("fake_devs", ctypes.c_int * 8))[: g.fake("fake_ndev").value]    # noqa: E731
    for o, want, created, destroyed in ((OPTS, [0], 1, 0), (dict(OPTS, ngpu=8.0), list(range(8)), 2, 1),
                                        (dict(OPTS, ngpu=8.0), list(range(8)), 2, 1),
                                        (dict(OPTS, devices=np.array([[1.0, 3.0]])), [1, 3], 3, 2), (OPTS, [0], 4, 3)):
        out, err = g.call(5, D, r, g.struct(o))
        assert err is None
        assert devs() == want and g.fake("fake_created").value == created and g.fake("fake_destroyed").value == destroyed
    assert g.lib.mock_locks() == 1
    assert g.lib.mock_run_atexit() == 1 and g.fake("fake_destroyed").value == 4


def test_admm_gateway_mask_option_and_library_errors(built):
    g = _admm(built)
    D = _problem()
    g.randn_source(1000)
    mask = np.random.default_rng(7).random(D.shape) > 0.3
    out, err = g.call(5, g.double(D), g.double(2), g.struct(dict(OPTS, mask=g.logical(mask))))
    assert err is None and g.fake("fake_mask_seen").value == 1 and g.fake("fake_mask_sum", ctypes.c_long).value == int(mask.sum())
    _, err = g.call(5, g.double(D), g.double(2), g.struct(dict(OPTS, mask=g.double(mask.astype(float)))))
    assert err == ("tritd:opts", "opts.mask must be a logical array of the size of D.")
    g.fake("fake_fail_solve").value = 4                                            # TRITD_ERR_NUMERIC from the solve
    out, err = g.call(5, g.double(D), g.double(2), g.struct(OPTS))
    assert out is None and err[0] == "tritd:solve" and "NaN / Inf" in err[1]       # tritd_last_error() is the message
    g2 = _admm(built)
    g2.randn_source(1000)
    g2.fake("fake_fail_create").value = 1
    out, err = g2.call(5, g2.double(D), g2.double(2), g2.struct(OPTS))
    assert out is None and err[0] == "tritd:cuda" and g2.lib.mock_locks() == 0     # nothing is locked when creation failed


def test_masked_gateway(built):
    sos, d = built
    g = Gateway(sos["triple_ADMM_masked"], d)
    n1, n2, n3, r = 6, 5, 4, 2
    D = _problem(n1, n2, n3)
    mask = np.random.default_rng(9).random(D.shape) > 0.25
    g.randn_source(1000)
    out, err = g.call(6, g.double(D), g.logical(mask), g.double(r), g.struct(OPTS))    # [A,B,C,O,E,Out]
    assert err is None
    assert g.randn_calls() == [(n1, r, r), (r, n2, r), (r, r, n3)]
    assert g.fake("fake_mask_sum", ctypes.c_long).value == int(mask.sum())
    assert np.array_equal(out[3], D + 1) and np.array_equal(out[4], D + 2)
    eh = g.field(out[5], "errHist")                                               # errHist = Out.errHist (traffic_triple_comparison.m:54)
    assert eh.shape == (7, 1)
    _, err = g.call(6, g.double(D), g.double(mask.astype(float)), g.double(r), g.struct(OPTS))
    assert err[0] == "tritd:mask"
    _, err = g.call(6, g.double(D), g.logical(mask[:, :, :3]), g.double(r), g.struct(OPTS))
    assert err[0] == "tritd:mask"
    _, err = g.call(7, g.double(D), g.logical(mask), g.double(r), g.struct(OPTS))
    assert err[0] == "MATLAB:TooManyOutputs"
    _, err = g.call(6, g.double(D), g.double(r), g.struct(OPTS))
    assert err[0] == "tritd:nargin"


def test_als_gateway(built):
    sos, d = built
    g = Gateway(sos["triple_decomp_ALS"], d)
    n1, n2, n3, r = 6, 5, 4, 3
    X = _problem(n1, n2, n3)
    g.randn_source(1000)
    out, err = g.call(4, g.double(X), g.double(r), g.struct(dict(maxIter=12, tol=1e-6)))     # the two fields the reference reads (:2-3)
    assert err is None
    assert g.randn_calls() == [(n1, r, r), (r, n2, r), (r, r, n3)]                            # triple_decomp_ALS.m:8-10
    assert out[0].shape == (n1, r, r) and out[1].shape == (r, n2, r) and out[2].shape == (r, r, n3) and out[3].shape == (7, 1)
    o = g.fake("fake_opts", _Opts)
    assert (o.maxIter, o.tol, o.disp) == (12, 1e-6, 1)                                        # the reference always prints (:17-19)
    assert g.printed().startswith("Iteration 5, relative error = ")
    out, err = g.call(4, g.double(X), g.double(r), g.struct(dict(maxIter=12, tol=1e-6, device=3.0)))     # another device: a new context
    assert err is None and g.fake("fake_created").value == 2 and g.fake("fake_destroyed").value == 1
    assert list(g.fake("fake_devs", ctypes.c_int * 8))[:1] == [3] and g.lib.mock_locks() == 1
    _, err = g.call(4, g.double(X), g.double(r), g.struct(dict(maxIter=12)))
    assert err == ("MATLAB:nonExistentField", 'Unrecognized field name "tol".')
    _, err = g.call(5, g.double(X), g.double(r), g.struct(dict(maxIter=12, tol=1e-6)))
    assert err[0] == "MATLAB:TooManyOutputs"


def test_triple_product_gateway(built):
    sos, d = built
    g = Gateway(sos["triple_product"], d)
    n1, n2, n3, r = 5, 4, 3, 2
    rng = np.random.default_rng(11)
    A, B, C = rng.standard_normal((n1, r, r)), rng.standard_normal((r, n2, r)), rng.standard_normal((r, r, n3))
    out, err = g.call(1, g.double(A), g.double(B), g.double(C))
    assert err is None and out[0].shape == (n1, n2, n3)
    assert np.allclose(out[0], orc.triple_product(A, B, C), rtol=1e-13, atol=1e-13)          # (the double computes the definition)
    assert list(g.fake("fake_shape", ctypes.c_longlong * 4)) == [n1, n2, n3, r]
    _, err = g.call(1, g.double(A), g.double(np.zeros((r + 1, n2, r))), g.double(C))
    assert err[0] == "tritd:arg"
    _, err = g.call(2, g.double(A), g.double(B), g.double(C))
    assert err[0] == "tritd:nargin"
    # C with n3 = 1 arrives as an r x r matrix
    out, err = g.call(1, g.double(A), g.double(B), g.double(C[:, :, 0]))
    assert err is None and out[0].shape == (n1, n2) and list(g.fake("fake_shape", ctypes.c_longlong * 4)) == [n1, n2, 1, r]


def test_gateway_on_the_real_library_reports_a_missing_device(tmp_path):
    """Linked against the real libtritd on a machine without a GPU: the gateway must fail with tritd:cuda (no CPU fallback)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present: tests/test_zz_mex_gateway_gpu.py runs the solve")
    except ImportError:
        pass
    so = build_gateway(tmp_path, "triple_decomp_ADMM", real=True)
    g = Gateway(so, tmp_path)
    g.randn_source(1000)
    out, err = g.call(5, g.double(_problem()), g.double(2), g.struct(OPTS))
    assert out is None and err[0] == "tritd:cuda" and "no CPU fallback" in err[1]
    assert g.randn_calls() == [(6, 2, 2), (2, 5, 2), (2, 2, 4)]                   # the factors had been drawn, as in the reference
