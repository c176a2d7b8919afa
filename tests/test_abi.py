"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/tritd.h declares, fails loudly without a GPU (no CPU fallback), and the host-side
mirror validates arguments like the reference does."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import tritd
from tritd import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tritd.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tritd_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    lib = tritd.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"libtritd.so does not export {name}"
    assert set(declared) == set(tritd.SYMBOLS), "ctypes table out of sync with include/tritd.h"


def test_opts_struct_layout_matches_header():
    assert ctypes.sizeof(tritd.tritd_opts) == 5 * 8 + 2 * 4
    assert ctypes.sizeof(tritd.tritd_timing) == 4 * 8 + 2 * 4
    assert tritd.tritd_opts.maxIter.offset == 40 and tritd.tritd_opts.disp.offset == 44


def test_version_and_slab_bounds_need_no_gpu():
    lib = tritd.load_library()
    assert b"sm_100a" in lib.tritd_version()
    for n3, nr in ((300, 8), (512, 8), (5, 2), (7, 7), (3, 1)):
        assert [tritd.slab_bounds(n3, nr, g) for g in range(nr)] == synth.slab_bounds(n3, nr)
    with pytest.raises(tritd.TritdError):
        tritd.slab_bounds(10, 2, 2)


def test_no_silent_cpu_fallback():
    """Without a CUDA device the library must refuse, not compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(tritd.TritdError) as ei:
        tritd.Context(0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_missing_opts_field_raises_like_matlab():
    o = dict(synth.TRAFFIC_OPTS)
    del o["rho"]
    with pytest.raises(KeyError, match='Unrecognized field name "rho"'):
        tritd.make_opts(o)
    o = dict(synth.TRAFFIC_OPTS, alphaA=1e-3, alphaB=1e-3, origin=np.zeros(3))   # set-but-ignored by the reference
    c = tritd.make_opts(o)
    assert c.maxIter == 100 and c.lambda_ == 1.8 and c.disp == 0


def test_argument_validation_before_any_gpu_work():
    with pytest.raises(ValueError):
        tritd.triple_decomp_ADMM(np.zeros((3, 3, 3, 3)), 2, synth.TRAFFIC_OPTS)
    with pytest.raises(TypeError):
        tritd.triple_decomp_ADMM(np.zeros((3, 3, 3), dtype=complex), 2, synth.TRAFFIC_OPTS)
    with pytest.raises(ValueError):
        tritd.triple_decomp_ADMM(np.zeros((3, 3, 3)), 0, synth.TRAFFIC_OPTS)
    with pytest.raises(ValueError, match="Mode must be 1, 2, or 3."):
        tritd.unfold(np.zeros((2, 2, 2)), 4)
    with pytest.raises(ValueError):
        tritd.buildF(np.zeros((2, 5, 3)), np.zeros((2, 2, 4)))


def test_mex_gateway_sources_present():
    mex = os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200", "mex")
    for f in ("triple_decomp_ADMM.c", "triple_product.c"):
        assert os.path.exists(os.path.join(mex, f))


def test_every_entry_point_rejects_a_null_handle():
    """No entry point that takes a context / problem handle may crash on NULL: each must return an error status (and none
    may compute anything without a device).  Run in a child process so that a crash fails this test, not the session."""
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
import tritd
lib = tritd.load_library()
n = 0
for name, (res, args) in tritd.SYMBOLS.items():
    if res is not C.c_int or not args or args[0] is not C.c_void_p:
        continue
    zero = [0 if a in (C.c_int, C.c_int32, C.c_int64) else 0.0 if a is C.c_double else None for a in args]
    rc = getattr(lib, name)(*zero)
    assert rc != 0, name
    assert lib.tritd_last_error(), name
    n += 1
print("rejected", n)
''' % os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert int(out.stdout.split()[-1]) >= 35
