"""Diagnostics: globaltimer stamps inside k_upd (A, B, C updates) of the last iteration."""
import os, sys, ctypes as C
os.environ["TRITD_DEBUG_STAMPS"] = "1"
os.environ["TRITD_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np, tritd
from tritd import synth
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(240, 320, 300, 5), (1024, 1024, 64, 8)]
for (n1, n2, n3, r) in shapes:
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    with tritd.Problem(tritd.default_context(), n1, n2, n3, r) as p:
        p.set_D(D); p.init(dict(synth.VIDEO_OPTS, maxIter=5000, tol=0.0), A0, B0, C0)
        p.enqueue(3000 if n1*n2*n3 < 5e7 else 600); p.sync()
        out = (C.c_longlong * 48)()
        lib = tritd.load_library()
        lib.tritd_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
        assert lib.tritd_debug_stamps(p._h, out) == 0
        for w, nm in enumerate("ABC"):
            s = list(out)[16 * w:16 * w + 16]
            b = s[0]
            print((n1, n2, n3, r), nm, "ns since block0 start: inv_init=%d inv_done=%d gram_wait_done=%d chunk_loaded=%d computed=%d end=%d | block1: start=%d reduced=%d inv_seen=%d Ms_loaded=%d applied=%d XT_written=%d rows_done=%d" % (
                s[13] - b, s[1] - b, s[2] - b, s[11] - b, s[12] - b, s[3] - b, s[4] - b, s[5] - b, s[6] - b, s[8] - b, s[9] - b, s[10] - b, s[7] - b))
