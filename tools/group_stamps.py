"""Diagnostics: k_upd stamps of the last iteration on every device of a single-process multi-GPU context."""
import os, sys
os.environ["TRITD_DEBUG_STAMPS"] = "1"
os.environ["TRITD_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np, tritd, torch
from tritd import synth
nd = torch.cuda.device_count()
for arg in sys.argv[1:]:
    n1, n2, n3, r = (int(x) for x in arg.split("x"))
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    with tritd.Context.from_devices(list(range(nd))) as g:
        print(arg, "on", nd, "devices", flush=True)
        out = tritd.triple_decomp_ADMM(D, r, dict(synth.VIDEO_OPTS, maxIter=300, tol=0.0), A0, B0, C0, ctx=g, return_info=True)
        print("  iterate_ms per iteration: %.1f us" % (out[5]["iterate_ms"] / 300 * 1e3), flush=True)
