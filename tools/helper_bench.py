"""GB/s of the standalone L2 helpers on device-resident data (north_star item 4: "coalesced, vectorised ... evidenced by
GB/s"): unfold (modes 2, 3), buildF/G/H, soft_threshold, triple_product -- the *_dev_f64 entry points timed with CUDA
events on the library's stream, cfg3 shape by default.  python tools/helper_bench.py [n1xn2xn3xr]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np
import torch
import tritd
from tritd import synth

n1, n2, n3, r = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "240x320x300x5").split("x"))
R, N = r * r, n1 * n2 * n3
lib = tritd.load_library()
ctx = tritd.default_context()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
A, B, Cc = synth.init_factors(n1, n2, n3, r, 1)
X = torch.randn(N, dtype=torch.float64, device="cuda")
Y = torch.empty(N, dtype=torch.float64, device="cuda")
dA = torch.from_numpy(np.ascontiguousarray(A.ravel(order="F"))).cuda()
dB = torch.from_numpy(np.ascontiguousarray(B.ravel(order="F"))).cuda()
dC = torch.from_numpy(np.ascontiguousarray(Cc.ravel(order="F"))).cuda()
F = torch.empty(R * max(n2 * n3, n1 * n3, n1 * n2), dtype=torch.float64, device="cuda")
vp = C.c_void_p


def timed(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms * 1e3:9.1f} us   {nbytes / ms * 1e-6:8.0f} GB/s  ({nbytes / 1e6:.0f} MB moved)", flush=True)


def chk(code):
    if code != 0:
        raise RuntimeError(lib.tritd_last_error().decode())


print(f"shape {n1}x{n2}x{n3}, r={r} (one N-array = {N * 8e-6:.0f} MB)")
for mode in (2, 3):
    timed(f"unfold(X,{mode})  [read N + write N]", lambda m=mode: chk(lib.tritd_unfold_dev_f64(ctx._h, vp(X.data_ptr()), n1, n2, n3, m, vp(Y.data_ptr()))), 16 * N)
timed("buildF(B,C)  [write R*n2*n3]", lambda: chk(lib.tritd_build_design_dev_f64(ctx._h, 0, vp(dB.data_ptr()), vp(dC.data_ptr()), n2, n3, r, vp(F.data_ptr()))), 8 * R * n2 * n3)
timed("buildG(A,C)  [write R*n1*n3]", lambda: chk(lib.tritd_build_design_dev_f64(ctx._h, 1, vp(dA.data_ptr()), vp(dC.data_ptr()), n1, n3, r, vp(F.data_ptr()))), 8 * R * n1 * n3)
timed("buildH(A,B)  [write R*n1*n2]", lambda: chk(lib.tritd_build_design_dev_f64(ctx._h, 2, vp(dA.data_ptr()), vp(dB.data_ptr()), n1, n2, r, vp(F.data_ptr()))), 8 * R * n1 * n2)
timed("soft_threshold(X,lam)  [read N + write N]", lambda: chk(lib.tritd_soft_threshold_dev_f64(ctx._h, vp(X.data_ptr()), N, 0.5, vp(Y.data_ptr()))), 16 * N)
Ah, Bh, Ch = (np.asfortranarray(x) for x in (A, B, Cc))
timed("triple_product(A,B,C) -> device  [write N; incl. factor upload + set-up]",
      lambda: chk(lib.tritd_triple_product_dev_f64(ctx._h, Ah.ctypes.data_as(vp), Bh.ctypes.data_as(vp), Ch.ctypes.data_as(vp), n1, n2, n3, r, vp(Y.data_ptr()))), 8 * N, reps=5)
# correctness spot checks against numpy on the same inputs
Xh = X.cpu().numpy().reshape((n1, n2, n3), order="F")
lib.tritd_unfold_dev_f64(ctx._h, vp(X.data_ptr()), n1, n2, n3, 3, vp(Y.data_ptr())); torch.cuda.synchronize()
assert np.array_equal(Y.cpu().numpy().reshape((n3, n1 * n2), order="F"), np.reshape(np.transpose(Xh, (2, 0, 1)), (n3, n1 * n2), order="F"))
print("unfold mode 3 bit-exact vs numpy: ok")
