"""Wall-clock time per iteration of the graph-replay path (what bench.py's `value` measures) for arbitrary shapes."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np
import tritd
from tritd import synth

def run(n1, n2, n3, r, iters=1000):
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    ctx = tritd.default_context()
    big = n1 * n2 * n3 >= 5e7
    warm, iters = (300, 200) if big else (2000, iters)
    with tritd.Problem(ctx, n1, n2, n3, r) as p:
        p.set_D(D)
        p.init(dict(synth.VIDEO_OPTS, maxIter=warm + 3 * iters + 10, tol=0.0), A0, B0, C0)
        p.enqueue(warm); p.sync()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); p.enqueue(iters); p.sync(); best = min(best, (time.perf_counter() - t0) / iters)
    print(f"{n1}x{n2}x{n3} r={r}: {best * 1e6:.1f} us/iteration ({1 / best:.0f} it/s)", flush=True)

if __name__ == "__main__":
    for a in sys.argv[1:] or ["240x320x300x5"]:
        run(*(int(x) for x in a.split("x")))
