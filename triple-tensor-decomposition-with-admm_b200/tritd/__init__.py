"""Host-side mirror of the reference's MATLAB interface over the libtritd C ABI.

The reference (dangnq2501/Triple-Tensor-Decomposition-with-ADMM) is pure MATLAB;
neither MATLAB nor Octave exists in this image, so the host side above the C ABI
is this thin ctypes layer with the reference's own names, argument meaning and
error behaviour:

    [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)   triple_decomp_ADMM.m:1
    Xhat = triple_product(A,B,C)                         triple_product.m:1
    Xn   = unfold(X, mode)                               unfold.m:1
    F/G/H = buildF(B,C) / buildG(A,C) / buildH(A,B)      buildF.m / buildG.m / buildH.m
    O    = soft_threshold(X, lam)                        soft_threshold.m:1

Arrays are numpy float64 in Fortran (column-major) order, shapes as in MATLAB.
Everything computes on the GPU through ``libtritd.so``; there is no CPU fallback
-- if the library or a CUDA device is missing the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (re-export)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtritd.so")

REQUIRED_OPTS = ("mu", "rho", "lambda", "lambda2", "maxIter", "tol", "disp")
MAX_R = 8
NCCL_ID_BYTES = 128
NPHASE = 6
PHASES = ("mttkrp1", "solveA", "ppass", "solveBC", "fused", "finalize")


class TritdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tritd error {code}: {msg}")
        self.code = code


class tritd_opts(C.Structure):
    _fields_ = [("mu", C.c_double), ("rho", C.c_double), ("lambda_", C.c_double), ("lambda2", C.c_double),
                ("tol", C.c_double), ("maxIter", C.c_int32), ("disp", C.c_int32)]


class tritd_timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("iterate_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("iters", C.c_int32), ("launches", C.c_int32)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64

# every symbol include/tritd.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "tritd_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "tritd_nccl_unique_id": (C.c_int, [_vp]),
    "tritd_create_rank": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "tritd_create_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "tritd_destroy": (None, [_vp]),
    "tritd_last_error": (C.c_char_p, []),
    "tritd_version": (C.c_char_p, []),
    "tritd_set_stream": (C.c_int, [_vp, _vp]),
    "tritd_slab_bounds": (C.c_int, [_i64, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64)]),
    "tritd_admm_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, C.c_int, C.POINTER(tritd_opts), _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int32), C.POINTER(tritd_timing)]),
    "tritd_admm_ex_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, C.c_int, C.POINTER(tritd_opts), _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int32), C.POINTER(tritd_timing)]),
    "tritd_set_print": (None, [_vp, _vp]),
    "tritd_trim": (C.c_int, [_vp]),
    "tritd_evaluate_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp, _vp, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "tritd_design_qi_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "tritd_triple_product_qi_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "tritd_als_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int32, C.c_double, C.c_int32, _vp, _vp, _vp,
                                _vp, _vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "tritd_problem_create": (C.c_int, [_vp, _i64, _i64, _i64, C.c_int, C.POINTER(_vp)]),
    "tritd_problem_destroy": (None, [_vp]),
    "tritd_problem_set_D_host": (C.c_int, [_vp, _vp]),
    "tritd_problem_set_D_dev": (C.c_int, [_vp, _vp]),
    "tritd_problem_set_mask_host": (C.c_int, [_vp, _vp]),
    "tritd_problem_set_mask_dev": (C.c_int, [_vp, _vp]),
    "tritd_problem_get_E": (C.c_int, [_vp, _vp]),
    "tritd_problem_get_E_dev": (C.c_int, [_vp, _vp]),
    "tritd_problem_evaluate": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tritd_problem_pinv_stats": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tritd_problem_init": (C.c_int, [_vp, C.POINTER(tritd_opts), _vp, _vp, _vp]),
    "tritd_problem_iterate": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_int32)]),
    "tritd_problem_enqueue": (C.c_int, [_vp, C.c_int32]),
    "tritd_problem_sync": (C.c_int, [_vp]),
    "tritd_problem_get": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "tritd_problem_get_O_dev": (C.c_int, [_vp, _vp]),
    "tritd_problem_get_L_dev": (C.c_int, [_vp, _vp]),
    "tritd_launch_count": (C.c_int64, [_vp]),
    "tritd_measure_dmma_peak": (C.c_int, [_vp, C.c_double, C.POINTER(C.c_double)]),
    "tritd_problem_set_profiling": (C.c_int, [_vp, C.c_int]),
    "tritd_problem_phase_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_int32)]),
    "tritd_triple_product_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "tritd_triple_product_dev_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "tritd_unfold_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "tritd_unfold_dev_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "tritd_build_design_dev_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "tritd_soft_threshold_dev_f64": (C.c_int, [_vp, _vp, _i64, C.c_double, _vp]),
    "tritd_factor_update_f64": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp]),
    "tritd_buildF_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "tritd_buildG_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "tritd_buildH_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i64, C.c_int, _vp]),
    "tritd_soft_threshold_f64": (C.c_int, [_vp, _vp, _i64, C.c_double, _vp]),
    "tritd_mttkrp_f64": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, C.c_int, C.c_int, _vp]),
}

_lib = None


def load_library():
    """dlopen libtritd.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(tritd has no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(code):
    if code != 0:
        raise TritdError(code, load_library().tritd_last_error().decode())


def _f64(a, shape=None):
    """float64 column-major view/copy of ``a``; MATLAB-style size check."""
    x = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and tuple(x.shape) != tuple(shape):
        raise ValueError(f"array of size {tuple(x.shape)} where {tuple(shape)} was expected")
    return x


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def make_opts(opts) -> tritd_opts:
    """opts struct -> C struct.  A missing field is a hard error, like MATLAB's
    'Unrecognized field name' at triple_decomp_ADMM.m:16-20; unknown fields
    (alphaA, alphaB, origin, ...) are ignored like the reference ignores them."""
    for k in REQUIRED_OPTS:
        if k not in opts:
            raise KeyError(f'Unrecognized field name "{k}".')
    return tritd_opts(float(opts["mu"]), float(opts["rho"]), float(opts["lambda"]), float(opts["lambda2"]),
                      float(opts["tol"]), int(opts["maxIter"]), int(bool(opts["disp"])))


class Context:
    """One GPU (optionally one rank of a mode-3 sharded solve)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None):
        lib = load_library()
        h = _vp()
        if nranks == 1:
            _check(lib.tritd_create(int(device), C.byref(h)))
        else:
            buf = C.create_string_buffer(bytes(nccl_id), NCCL_ID_BYTES)
            _check(lib.tritd_create_rank(int(device), int(rank), int(nranks), buf, C.byref(h)))
        self._h = h
        self.device, self.rank, self.nranks = device, rank, nranks

    @classmethod
    def from_devices(cls, devices):
        """Single-process multi-GPU context (tritd_create_devices): full tensors in and out of triple_decomp_ADMM."""
        devices = [int(d) for d in devices]
        self = cls.__new__(cls)
        h = _vp()
        arr = (C.c_int * len(devices))(*devices)
        _check(load_library().tritd_create_devices(arr, len(devices), C.byref(h)))
        self._h = h
        self.device, self.rank, self.nranks = devices[0], 0, 1
        self.devices = devices
        return self

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(NCCL_ID_BYTES)
        _check(load_library().tritd_nccl_unique_id(buf))
        return buf.raw

    def set_stream(self, cuda_stream_ptr):
        _check(load_library().tritd_set_stream(self._h, _vp(cuda_stream_ptr or 0)))

    @property
    def launches(self) -> int:
        return int(load_library().tritd_launch_count(self._h))

    def measure_dmma_peak(self, ms_budget=50.0) -> float:
        """measured FP64 tensor-core (DMMA) peak of this GPU, TFLOP/s"""
        t = C.c_double()
        _check(load_library().tritd_measure_dmma_peak(self._h, float(ms_budget), C.byref(t)))
        return t.value

    def trim(self):
        """Release the device state tritd_admm_f64 caches between equally shaped calls."""
        _check(load_library().tritd_trim(self._h))

    def close(self):
        if self._h:
            load_library().tritd_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("TRITD_DEVICE", "0")))
    return _default_ctx


def slab_bounds(n3, nranks, rank):
    t0, t1 = _i64(), _i64()
    _check(load_library().tritd_slab_bounds(int(n3), int(nranks), int(rank), C.byref(t0), C.byref(t1)))
    return t0.value, t1.value


class Problem:
    """Device-resident state of one solve (the staged half of the C ABI)."""

    def __init__(self, ctx: Context, n1, n2, n3_local, r):
        self.ctx = ctx
        self.shape = (int(n1), int(n2), int(n3_local))
        self.r = int(r)
        h = _vp()
        _check(load_library().tritd_problem_create(ctx._h, *self.shape, self.r, C.byref(h)))
        self._h = h
        self.max_iter = 0

    def set_D(self, D):
        D = _f64(D, self.shape)
        _check(load_library().tritd_problem_set_D_host(self._h, _ptr(D)))

    def set_D_dev(self, dev_ptr):
        _check(load_library().tritd_problem_set_D_dev(self._h, _vp(dev_ptr)))

    def set_mask(self, mask):
        """completion variant: boolean tensor of D's shape, True = observed (call after set_D, before init)"""
        m = np.asfortranarray(np.asarray(mask).astype(bool), dtype=np.uint8)
        if m.shape != self.shape:
            raise ValueError("mask must have the shape of D")
        _check(load_library().tritd_problem_set_mask_host(self._h, m.ctypes.data_as(_vp)))

    def get_E(self):
        E = np.zeros(self.shape, order="F")
        _check(load_library().tritd_problem_get_E(self._h, _ptr(E)))
        return E

    def evaluate(self, gt, mask=None):
        """[rmse, nrmse] = evaluate(triple_product(A,B,C), gt, mask) with the resident factors"""
        gt = _f64(gt, self.shape)
        m = None
        if mask is not None:
            m = np.asfortranarray(np.asarray(mask).astype(bool), dtype=np.uint8)
        rmse, nrmse = C.c_double(), C.c_double()
        _check(load_library().tritd_problem_evaluate(self._h, _ptr(gt), m.ctypes.data_as(_vp) if m is not None else None,
                                                     C.byref(rmse), C.byref(nrmse)))
        return rmse.value, nrmse.value

    def pinv_stats(self):
        """(ridge solves that took the truncating pinv path, singular values they zeroed)"""
        a, b = C.c_int32(), C.c_int32()
        _check(load_library().tritd_problem_pinv_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def init(self, opts, A0, B0, C0):
        n1, n2, n3 = self.shape
        r = self.r
        A0 = _f64(A0, (n1, r, r)); B0 = _f64(B0, (r, n2, r)); C0 = _f64(C0, (r, r, n3))
        o = make_opts(opts)
        self.max_iter = o.maxIter
        _check(load_library().tritd_problem_init(self._h, C.byref(o), _ptr(A0), _ptr(B0), _ptr(C0)))

    def iterate(self, max_more=None) -> int:
        k = C.c_int32()
        _check(load_library().tritd_problem_iterate(self._h, int(self.max_iter if max_more is None else max_more),
                                                    C.byref(k)))
        return k.value

    def enqueue(self, n):
        _check(load_library().tritd_problem_enqueue(self._h, int(n)))

    def sync(self):
        _check(load_library().tritd_problem_sync(self._h))

    def get(self, want_O=True, want_L=False):
        n1, n2, n3 = self.shape
        r = self.r
        A = np.zeros((n1, r, r), order="F"); B = np.zeros((r, n2, r), order="F"); Cc = np.zeros((r, r, n3), order="F")
        O = np.zeros(self.shape, order="F") if want_O else None
        L = np.zeros(self.shape, order="F") if want_L else None
        eh = np.zeros(self.max_iter); eL = np.zeros(self.max_iter); eO = np.zeros(self.max_iter)
        k = C.c_int32()
        _check(load_library().tritd_problem_get(self._h, _ptr(A), _ptr(B), _ptr(Cc), _ptr(O), _ptr(L), _ptr(eh),
                                                _ptr(eL), _ptr(eO), C.byref(k)))
        k = k.value
        return dict(A=A, B=B, C=Cc, O=O, L=L, errHist=eh[:k].copy(), errL=eL[:k].copy(), errO=eO[:k].copy(), iters=k)

    def set_profiling(self, enable=True):
        _check(load_library().tritd_problem_set_profiling(self._h, int(bool(enable))))

    def phase_ms(self):
        """(ms per phase summed over the recorded iterations, iterations recorded); see TRITD_NPHASE in tritd.h."""
        out = (C.c_double * NPHASE)()
        n = C.c_int32()
        _check(load_library().tritd_problem_phase_ms(self._h, out, C.byref(n)))
        return list(out), n.value

    def get_O_dev(self, dev_ptr):
        _check(load_library().tritd_problem_get_O_dev(self._h, _vp(dev_ptr)))

    def get_L_dev(self, dev_ptr):
        _check(load_library().tritd_problem_get_L_dev(self._h, _vp(dev_ptr)))

    def close(self):
        if self._h:
            load_library().tritd_problem_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# --------------------------------------------------------------------------
# the reference's functions
# --------------------------------------------------------------------------
def triple_decomp_ADMM(D, r, opts, A0=None, B0=None, C0=None, rng=None, ctx=None, return_info=False, want_L=False,
                       out_O=None, mask=None, want_E=False):
    """[A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)   (triple_decomp_ADMM.m:1-70).

    The reference draws A,B,C with randn at :23 (order A, B, C).  Here the same
    order is drawn from ``rng`` (numpy Generator; default_rng(0) if omitted)
    unless ``A0,B0,C0`` (or ``opts['A0']`` ...) inject them -- the library itself
    never draws random numbers.  ``out_O`` may be a preallocated Fortran-ordered float64 array (e.g. a view of pinned
    memory) to receive O.  Extensions (default path untouched): ``want_E`` adds E to ``info``;
    ``mask`` (boolean, True = observed; also ``opts['mask']``) selects the completion variant, see
    :func:`triple_ADMM_masked`."""
    D = np.asarray(D)
    if D.ndim == 2:                       # MATLAB hands an n1 x n2 x 1 tensor over as 2-D
        D = D[:, :, None]
    if D.ndim != 3:
        raise ValueError("D must be a 3-D array")
    if np.iscomplexobj(D):
        raise TypeError("D must be real")
    r = int(r)
    if r < 1:
        raise ValueError("r must be a positive integer")
    o = make_opts(opts)
    D = _f64(D)
    n1, n2, n3 = D.shape
    A0 = opts.get("A0", A0); B0 = opts.get("B0", B0); C0 = opts.get("C0", C0)
    if A0 is None or B0 is None or C0 is None:
        rng = rng or np.random.default_rng(0)
        A0 = rng.standard_normal((n1, r, r)); B0 = rng.standard_normal((r, n2, r)); C0 = rng.standard_normal((r, r, n3))
    A0 = _f64(A0, (n1, r, r)); B0 = _f64(B0, (r, n2, r)); C0 = _f64(C0, (r, r, n3))
    ctx = ctx or default_context()
    A = np.zeros((n1, r, r), order="F"); B = np.zeros((r, n2, r), order="F"); Cc = np.zeros((r, r, n3), order="F")
    if out_O is not None:
        if out_O.shape != (n1, n2, n3) or out_O.dtype != np.float64 or not out_O.flags.f_contiguous:
            raise ValueError("out_O must be a Fortran-ordered float64 array of D's shape")
        O = out_O
    else:
        O = np.zeros((n1, n2, n3), order="F")
    L = np.zeros((n1, n2, n3), order="F") if want_L else None
    E = np.zeros((n1, n2, n3), order="F") if want_E else None
    mask = opts.get("mask", mask)
    m = None
    if mask is not None:
        m = np.asfortranarray(np.asarray(mask).astype(bool), dtype=np.uint8)
        if m.ndim == 2:
            m = m[:, :, None]
        if m.shape != (n1, n2, n3):
            raise ValueError("mask must have the shape of D")
    errHist = np.zeros(o.maxIter)
    k = C.c_int32()
    tm = tritd_timing()
    if m is None and E is None:          # the reference's own call: the 3-in / 5-out entry point
        _check(load_library().tritd_admm_f64(ctx._h, _ptr(D), n1, n2, n3, r, C.byref(o), _ptr(A0), _ptr(B0), _ptr(C0),
                                             _ptr(A), _ptr(B), _ptr(Cc), _ptr(O), _ptr(L), _ptr(errHist), C.byref(k),
                                             C.byref(tm)))
    else:
        _check(load_library().tritd_admm_ex_f64(ctx._h, _ptr(D), m.ctypes.data_as(_vp) if m is not None else None, n1, n2, n3,
                                                r, C.byref(o), _ptr(A0), _ptr(B0), _ptr(C0), _ptr(A), _ptr(B), _ptr(Cc),
                                                _ptr(O), _ptr(E), _ptr(L), _ptr(errHist), C.byref(k), C.byref(tm)))
    errHist = errHist[:k.value].copy()
    if return_info:
        info = dict(iters=k.value, h2d_ms=tm.h2d_ms, iterate_ms=tm.iterate_ms, d2h_ms=tm.d2h_ms, total_ms=tm.total_ms,
                    launches=tm.launches, L=L, E=E)
        return A, B, Cc, O, errHist, info
    return A, B, Cc, O, errHist


def triple_ADMM_masked(Y, mask, r, opts, A0=None, B0=None, C0=None, rng=None, ctx=None):
    """[A,B,C,O,E,Out] = triple_ADMM_masked(Y, mask, r, opts) -- the completion variant the reference's drivers name
    in a comment (traffic_triple_comparison.m:53, video_triple_comparison.m:52: ``triple_ADMM_masked(Y, ~mask_missing,
    r, opts)``, ``errHist = Out.errHist``) but do not ship.  ``mask`` True = observed.  Semantics: DESIGN.md 4.6."""
    A, B, Cc, O, eh, info = triple_decomp_ADMM(Y, r, opts, A0, B0, C0, rng=rng, ctx=ctx, return_info=True, mask=mask,
                                                want_E=True)
    return A, B, Cc, O, info["E"], dict(errHist=eh, iters=info["iters"])


_print_cb = None


def set_print(fn):
    """Route the progress lines through ``fn(str)`` (None: back to stdout) -- what a MEX gateway does with mexPrintf."""
    global _print_cb
    lib = load_library()
    if fn is None:
        _print_cb = None
        lib.tritd_set_print(None, None)
        return
    _print_cb = C.CFUNCTYPE(None, C.c_char_p, _vp)(lambda line, user: fn(line.decode()))
    lib.tritd_set_print(C.cast(_print_cb, _vp), None)


def factor_update(rhs, S1, S2, alpha, ctx=None):
    """X = rhs * pinv(S1 .* S2 + alpha*eye) through the solver's update kernel (update_A/B/C, :77-78/:86/:93).
    -> X, pinv, X'X, (took the truncating path?, singular values zeroed)"""
    rhs = _f64(rhs)
    n, R = rhs.shape
    r = int(round(R ** 0.5))
    if r * r != R:
        raise ValueError("rhs must have r^2 columns")
    S1 = _f64(S1, (R, R)); S2 = _f64(S2, (R, R))
    X = np.zeros((n, R), order="F"); Gi = np.zeros((R, R), order="F"); XtX = np.zeros((R, R), order="F")
    info = (C.c_int32 * 2)()
    ctx = ctx or default_context()
    _check(load_library().tritd_factor_update_f64(ctx._h, _ptr(rhs), n, r, _ptr(S1), _ptr(S2), float(alpha), _ptr(X), _ptr(Gi),
                                                  _ptr(XtX), info))
    return X, Gi, XtX, (int(info[0]), int(info[1]))


def triple_decomp_ALS(X, r, opts, A0=None, B0=None, C0=None, rng=None, ctx=None, disp=None):
    """[A,B,C,errHist] = triple_decomp_ALS(X, r, opts)   (triple_decomp_ALS.m:1-40).

    ``opts`` needs ``maxIter`` and ``tol`` (:2-3; a missing one raises the reference's
    ``Unrecognized field name``).  The reference prints its progress line every 5th iteration
    unconditionally (:17-19); pass ``disp=False`` (or ``opts['disp'] = 0``) to silence it.
    Initial factors: as in :func:`triple_decomp_ADMM`."""
    X = np.asarray(X)
    if X.ndim == 2:
        X = X[:, :, None]
    if X.ndim != 3:
        raise ValueError("X must be a 3-D array")
    if np.iscomplexobj(X):
        raise TypeError("X must be real")
    r = int(r)
    if r < 1:
        raise ValueError("r must be a positive integer")
    for f in ("maxIter", "tol"):
        if f not in opts:
            raise KeyError('Unrecognized field name "%s".' % f)
    maxIter, tol = int(opts["maxIter"]), float(opts["tol"])
    if disp is None:
        disp = bool(opts.get("disp", 1))
    X = _f64(X)
    n1, n2, n3 = X.shape
    A0 = opts.get("A0", A0); B0 = opts.get("B0", B0); C0 = opts.get("C0", C0)
    if A0 is None or B0 is None or C0 is None:
        rng = rng or np.random.default_rng(0)
        A0 = rng.standard_normal((n1, r, r)); B0 = rng.standard_normal((r, n2, r)); C0 = rng.standard_normal((r, r, n3))
    A0 = _f64(A0, (n1, r, r)); B0 = _f64(B0, (r, n2, r)); C0 = _f64(C0, (r, r, n3))
    ctx = ctx or default_context()
    A = np.zeros((n1, r, r), order="F"); B = np.zeros((r, n2, r), order="F"); Cc = np.zeros((r, r, n3), order="F")
    errHist = np.zeros(max(maxIter, 1))
    k = C.c_int32()
    _check(load_library().tritd_als_f64(ctx._h, _ptr(X), n1, n2, n3, r, maxIter, tol, int(bool(disp)), _ptr(A0), _ptr(B0),
                                        _ptr(C0), _ptr(A), _ptr(B), _ptr(Cc), _ptr(errHist), C.byref(k)))
    return A, B, Cc, errHist[:k.value].copy()


def _factor_dims(A=None, B=None, Cc=None):
    r = None
    if A is not None:
        A = _f64(A)
        if A.ndim != 3 or A.shape[1] != A.shape[2]:
            raise ValueError("A must be n1 x r x r")
        r = A.shape[1]
    if B is not None:
        B = _f64(B)
        if B.ndim != 3 or B.shape[0] != B.shape[2] or (r is not None and B.shape[0] != r):
            raise ValueError("B must be r x n2 x r")
        r = B.shape[0]
    if Cc is not None:
        Cc = _f64(Cc)
        if Cc.ndim != 3 or Cc.shape[0] != Cc.shape[1] or (r is not None and Cc.shape[0] != r):
            raise ValueError("C must be r x r x n3")
        r = Cc.shape[0]
    return A, B, Cc, r


def triple_product(A, B, Cc, ctx=None):
    """Xhat = triple_product(A,B,C)   (triple_product.m:1-8)."""
    A, B, Cc, r = _factor_dims(A, B, Cc)
    n1, n2, n3 = A.shape[0], B.shape[1], Cc.shape[2]
    X = np.zeros((n1, n2, n3), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_triple_product_f64(ctx._h, _ptr(A), _ptr(B), _ptr(Cc), n1, n2, n3, r, _ptr(X)))
    return X


def _design_qi(which, U, V, na, nb, r, ctx):
    out = np.zeros((r * r, na * nb), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_design_qi_f64(ctx._h, which, _ptr(U), _ptr(V), na, nb, r, _ptr(out)))
    return out


def buildF_qi(B, Cc, ctx=None):
    """F = buildF(B,C) of the original (Qi) model   (origin_triple_tensor/buildF.m:4-6)."""
    _, B, Cc, r = _factor_dims(None, B, Cc)
    return _design_qi(0, B, Cc, B.shape[1], Cc.shape[2], r, ctx)


def buildG_qi(A, Cc, ctx=None):
    """G = buildG(A,C) of the original (Qi) model   (origin_triple_tensor/buildG.m:9-11)."""
    A, _, Cc, r = _factor_dims(A, None, Cc)
    return _design_qi(1, A, Cc, A.shape[0], Cc.shape[2], r, ctx)


def buildH_qi(A, B, ctx=None):
    """H = buildH(A,B) of the original (Qi) model   (origin_triple_tensor/buildH.m:9-11)."""
    A, B, _, r = _factor_dims(A, B, None)
    return _design_qi(2, A, B, A.shape[0], B.shape[1], r, ctx)


def triple_product_qi(A, B, Cc, ctx=None):
    """Xhat(i,j,t) = sum_{p,q,s} A(i,q,s) B(p,j,s) C(p,q,t)   (origin_triple_tensor/triple_product.m)."""
    A, B, Cc, r = _factor_dims(A, B, Cc)
    n1, n2, n3 = A.shape[0], B.shape[1], Cc.shape[2]
    X = np.zeros((n1, n2, n3), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_triple_product_qi_f64(ctx._h, _ptr(A), _ptr(B), _ptr(Cc), n1, n2, n3, r, _ptr(X)))
    return X


def evaluate(A, B, Cc, gt, mask=None, ctx=None):
    """[rmse, nrmse] = evaluate(triple_product(A,B,C), gt(mask), mask)   (traffic_triple_comparison.m:194-202), with the
    reconstruction formed and compared on the device.  ``gt`` is the dense ground-truth tensor, ``mask`` a boolean
    tensor of the same shape (None = every entry, i.e. the drivers' RRE)."""
    A, B, Cc, r = _factor_dims(A, B, Cc)
    n1, n2, n3 = A.shape[0], B.shape[1], Cc.shape[2]
    gt = _f64(gt, (n1, n2, n3))
    m = None
    if mask is not None:
        m = np.asfortranarray(np.asarray(mask).astype(bool), dtype=np.uint8)
        if m.shape != (n1, n2, n3):
            raise ValueError("mask must have the shape of gt")
    rmse, nrmse = C.c_double(), C.c_double()
    ctx = ctx or default_context()
    _check(load_library().tritd_evaluate_f64(ctx._h, _ptr(A), _ptr(B), _ptr(Cc), n1, n2, n3, r, _ptr(gt),
                                             m.ctypes.data_as(_vp) if m is not None else None, C.byref(rmse), C.byref(nrmse)))
    return rmse.value, nrmse.value


def unfold(X, mode, ctx=None):
    """Xn = unfold(X, mode)   (unfold.m:1-14); mode is 1-based."""
    X = _f64(X)
    if X.ndim != 3:
        raise ValueError("X must be a 3-D array")
    if mode not in (1, 2, 3):
        raise ValueError("Mode must be 1, 2, or 3.")
    n1, n2, n3 = X.shape
    shape = {1: (n1, n2 * n3), 2: (n2, n1 * n3), 3: (n3, n1 * n2)}[mode]
    out = np.zeros(shape, order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_unfold_f64(ctx._h, _ptr(X), n1, n2, n3, int(mode), _ptr(out)))
    return out


def buildF(B, Cc, ctx=None):
    """F = buildF(B,C): r^2 x (n2 n3)   (buildF.m:17-21)."""
    _, B, Cc, r = _factor_dims(None, B, Cc)
    n2, n3 = B.shape[1], Cc.shape[2]
    F = np.zeros((r * r, n2 * n3), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_buildF_f64(ctx._h, _ptr(B), _ptr(Cc), n2, n3, r, _ptr(F)))
    return F


def buildG(A, Cc, ctx=None):
    """G = buildG(A,C): r^2 x (n1 n3)   (buildG.m:17-21)."""
    A, _, Cc, r = _factor_dims(A, None, Cc)
    n1, n3 = A.shape[0], Cc.shape[2]
    G = np.zeros((r * r, n1 * n3), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_buildG_f64(ctx._h, _ptr(A), _ptr(Cc), n1, n3, r, _ptr(G)))
    return G


def buildH(A, B, ctx=None):
    """H = buildH(A,B): r^2 x (n1 n2)   (buildH.m:17-21)."""
    A, B, _, r = _factor_dims(A, B, None)
    n1, n2 = A.shape[0], B.shape[1]
    H = np.zeros((r * r, n1 * n2), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_buildH_f64(ctx._h, _ptr(A), _ptr(B), n1, n2, r, _ptr(H)))
    return H


def soft_threshold(X, lam, ctx=None):
    """O = soft_threshold(X, lam)   (soft_threshold.m:2)."""
    X = _f64(X)
    out = np.zeros(X.shape, order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_soft_threshold_f64(ctx._h, _ptr(X), X.size, float(lam), _ptr(out)))
    return out


def mttkrp(X, A, B, Cc, mode, ctx=None):
    """X_(mode) * M' for M = F, G, H at the given factors (the dgemm of update_A/B/C, :78/:86/:93)."""
    X = _f64(X)
    A, B, Cc, r = _factor_dims(A, B, Cc)
    n1, n2, n3 = X.shape
    n = (n1, n2, n3)[mode - 1]
    out = np.zeros((n, r * r), order="F")
    ctx = ctx or default_context()
    _check(load_library().tritd_mttkrp_f64(ctx._h, _ptr(X), _ptr(A), _ptr(B), _ptr(Cc), n1, n2, n3, r, int(mode),
                                           _ptr(out)))
    return out
