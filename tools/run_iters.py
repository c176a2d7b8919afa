"""Run a few plain (non-graph) iterations of a shape; target for ncu captures."""
import os, sys
os.environ.setdefault("TRITD_NO_GRAPH", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np, tritd
from tritd import synth
n1, n2, n3, r = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "240x320x300x5").split("x"))
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rng = np.random.default_rng(0)
D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
with tritd.Problem(tritd.default_context(), n1, n2, n3, r) as p:
    p.set_D(D); p.init(dict(synth.VIDEO_OPTS, maxIter=iters + 5, tol=0.0), A0, B0, C0)
    p.enqueue(iters); p.sync()
    print("ok", p.get(want_O=False)["iters"])
