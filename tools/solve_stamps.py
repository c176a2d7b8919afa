import os, sys, ctypes as C
os.environ["TRITD_DEBUG_STAMPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np, tritd
from tritd import synth
for (n1, n2, n3, r) in [(240, 320, 30, 5), (512, 512, 16, 8)]:
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    with tritd.Problem(tritd.default_context(), n1, n2, n3, r) as p:
        p.set_D(D); p.init(dict(synth.VIDEO_OPTS, maxIter=10, tol=0.0), A0, B0, C0)
        p.enqueue(3); p.sync()
        out = (C.c_longlong * 8)()
        lib = tritd.load_library()
        lib.tritd_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
        assert lib.tritd_debug_stamps(p._h, out) == 0
        s = list(out)
        print((n1, n2, n3, r), "cycles: load=%d gj=%d apply=%d write=%d gram=%d fence+sync=%d tail=%d total=%d" % (
            s[1]-s[0], s[2]-s[1], s[3]-s[2], s[4]-s[3], s[5]-s[4], s[6]-s[5], s[7]-s[6], s[7]-s[0]))
