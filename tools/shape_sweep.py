"""Per-phase timing of the iteration for arbitrary shapes (random data); diagnostic tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import numpy as np
import tritd
from tritd import synth

def run(n1, n2, n3, r, iters=100):
    rng = np.random.default_rng(0)
    D = np.asfortranarray(rng.standard_normal((n1, n2, n3)))
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 1)
    ctx = tritd.default_context()
    with tritd.Problem(ctx, n1, n2, n3, r) as p:
        p.set_D(D)
        p.init(dict(synth.VIDEO_OPTS, maxIter=iters + 2100, tol=0.0), A0, B0, C0)
        p.enqueue(2000 if n1 * n2 * n3 < 5e7 else 300); p.sync()      # long warm-up: SM clocks ramp up slowly
        p.set_profiling(True)
        p.enqueue(iters)
        ms, n = p.phase_ms()
        p.set_profiling(False)
    N = n1 * n2 * n3
    per = [m / n for m in ms]
    print(f"{n1}x{n2}x{n3} r={r}: " + " ".join(f"{nm}={v*1e3:.1f}us" for nm, v in zip(tritd.PHASES, per)) +
          f" | total={sum(per)*1e3:.1f}us fused={48*N/per[4]*1e-6:.0f}GB/s ppass={8*N/per[2]*1e-6:.0f}GB/s", flush=True)

if __name__ == "__main__":
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(240, 320, 300, 5)]
    for s in shapes:
        run(*s)
