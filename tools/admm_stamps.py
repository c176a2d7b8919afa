"""Diagnostics: per-CTA start / end stamps of the last k_admm launch, summarised per i-tile."""
import os, sys, ctypes as C
os.environ["TRITD_DEBUG_STAMPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np, tritd
from tritd import synth
for arg in sys.argv[1:] or ["240x320x300x5"]:
    n1, n2, n3, r = (int(x) for x in arg.split("x"))
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    with tritd.Problem(tritd.default_context(), n1, n2, n3, r) as p:
        p.set_D(D); p.init(dict(synth.VIDEO_OPTS, maxIter=3000, tol=0.0), A0, B0, C0)
        p.enqueue(1500 if n1 * n2 * n3 < 5e7 else 200); p.sync()
        lib = tritd.load_library()
        out = (C.c_longlong * 2048)(); tab = (C.c_int * 3072)()
        lib.tritd_debug_admm_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n = lib.tritd_debug_admm_stamps(p._h, out, tab, 1024)
        st = np.array(out[:2 * n]).reshape(n, 2); tb = np.array(tab[:3 * n]).reshape(n, 3)
        t0 = st[:, 0].min()
        print(arg, "grid", n, "kernel span %.1f us" % ((st[:, 1].max() - t0) / 1e3))
        dur = (st[:, 1] - st[:, 0]) / 1e3
        top = np.argsort(-dur)[:5]
        print("  slowest CTAs (id, tile, slot, us):", [(int(i), int(tb[i, 0]), int(tb[i, 1]), round(float(dur[i]), 1)) for i in top], " CTA 0: %.1f us" % dur[0])
        for q in sorted(set(tb[:, 0])):
            m = tb[:, 0] == q
            d = (st[m, 1] - st[m, 0]) / 1e3
            print("  tile %d: %3d CTAs  start %.1f..%.1f us  duration min %.1f  mean %.1f  max %.1f us  end max %.1f" %
                  (q, m.sum(), (st[m, 0].min() - t0) / 1e3, (st[m, 0].max() - t0) / 1e3, d.min(), d.mean(), d.max(), (st[m, 1].max() - t0) / 1e3))
