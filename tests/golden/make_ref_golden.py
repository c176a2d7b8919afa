"""Generates tests/golden/ref_m/*.npz by EXECUTING THE REFERENCE'S OWN, UNMODIFIED .m SOURCES where they lie under
/root/reference with the MATLAB-subset interpreter oracle/mlab.py (no MATLAB / Octave exists in this image).

    python tests/golden/make_ref_golden.py          # needs /root/reference; run in the build container only

What is pinned by these fixtures: statement order, index conventions, reshape / permute orders, operator
association, the function-resolution order (local functions of triple_decomp_ADMM.m shadow the files on the path)
and the printed progress lines of
    fast_robust_triple_tensor/triple_decomp_ADMM.m   (randn at :23 shadowed so that A0, B0, C0 are injected)
    fast_robust_triple_tensor/triple_decomp_ALS.m
    fast_robust_triple_tensor/{triple_product,unfold,buildF,buildG,buildH,soft_threshold}.m
    origin_triple_tensor/{buildF,buildG,buildH,kronF,kronG,kronH}.m            (the Qi model)
    traffic_triple_comparison.m:194-202 (local function evaluate)
What is NOT pinned: MathWorks' built-ins (pinv / mtimes / norm are LAPACK / OpenBLAS through numpy here).
Inputs are regenerated from seeds (tritd.synth, tests/golden/make_golden.py); only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402
import mlab  # noqa: E402
from tritd import synth  # noqa: E402

REF = "/root/reference"
FR = os.path.join(REF, "fast_robust_triple_tensor")
OR = os.path.join(REF, "origin_triple_tensor")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_m")

# ADMM cases: the oracle-golden cases of make_golden.py (same inputs) + BASELINE config 1 at full size with the
# caller's own options (traffic_triple_comparison.m:42-51: maxIter 100, tol 1e-5, disp 1)
ADMM_CASES = dict(make_golden.CASES)
ADMM_CASES["cfg1_50x50x50_r5_full"] = ("cfg1", (50, 50, 50), 5, 100, 1e-5)
ALS_CASES = dict(make_golden.ALS_CASES)


def case_inputs(name):
    if name == "cfg1_50x50x50_r5_full":
        w = synth.make_config("cfg1")
        return w["D"], w["r"], dict(w["opts"], disp=1), w["A0"], w["B0"], w["C0"]
    D, r, o, A0, B0, C0 = make_golden.case_inputs(name)
    return D, r, dict(o, disp=1), A0, B0, C0


def randn_injector(*factors):
    """shadow of randn(n1,r,r) / randn(r,n2,r) / randn(r,r,n3) (triple_decomp_ADMM.m:23): hands out A0, B0, C0 in
    call order and checks the requested sizes -- the order and the sizes are part of what is pinned"""
    queue = list(factors)

    def randn(interp, args, nargout):
        a = queue.pop(0)
        want = tuple(int(mlab.scalar(x)) for x in args)
        assert want == a.shape, (want, a.shape)
        return [mlab.mat(a)]
    return randn


def run_admm(D, r, o, A0, B0, C0):
    outs, text, it = mlab.run_function([FR, OR], "triple_decomp_ADMM", [D, float(r), o], nargout=5,
                                       overrides={"randn": randn_injector(A0, B0, C0)})
    A, B, C, O, eh = outs
    n1, n2, n3 = D.shape
    return (A.reshape((n1, r, r), order="F"), B.reshape((r, n2, r), order="F"), C.reshape((r, r, n3), order="F"),
            O.reshape(D.shape, order="F"), eh.reshape(-1), text, it)


def run_als(X, r, o, A0, B0, C0):
    outs, text, it = mlab.run_function([FR, OR], "triple_decomp_ALS", [X, float(r), {"maxIter": float(o["maxIter"]), "tol": o["tol"]}],
                                       nargout=4, overrides={"randn": randn_injector(A0, B0, C0)})
    A, B, C, eh = outs
    n1, n2, n3 = X.shape
    return (A.reshape((n1, r, r), order="F"), B.reshape((r, n2, r), order="F"), C.reshape((r, r, n3), order="F"),
            eh.reshape(-1), text)


def helper_outputs():
    """the L2 helper files on small random inputs (n = (7,6,5), r = 3, all sizes distinct)"""
    n1, n2, n3, r = 7, 6, 5, 3
    A, B, C = synth.init_factors(n1, n2, n3, r, 7)
    X = np.asfortranarray(np.random.Generator(np.random.PCG64(8)).standard_normal((n1, n2, n3)))
    f = lambda d, name, args, nout=1: mlab.run_function([d], name, args, nargout=nout)[0]   # noqa: E731
    out = {}
    for mode in (1, 2, 3):
        out[f"unfold{mode}"] = f(FR, "unfold", [X, float(mode)])[0]
    out["buildF"] = f(FR, "buildF", [B, C])[0]
    out["buildG"] = f(FR, "buildG", [A, C])[0]
    out["buildH"] = f(FR, "buildH", [A, B])[0]
    out["triple_product"] = f(FR, "triple_product", [A, B, C])[0].reshape((n1, n2, n3), order="F")
    out["soft_threshold"] = f(FR, "soft_threshold", [X, 0.7])[0].reshape((n1, n2, n3), order="F")
    # the Qi model: RPAS form (buildF/G/H) and the Kronecker form it replaces (kronF/G/H) must agree
    out["qi_buildF"] = f(OR, "buildF", [B, C])[0]
    out["qi_buildG"] = f(OR, "buildG", [A, C])[0]
    out["qi_buildH"] = f(OR, "buildH", [A, B])[0]
    out["qi_kronF"] = f(OR, "kronF", [B, C])[0]
    # evaluate() of the traffic driver (a local function of a script file)
    gt = np.asfortranarray(np.random.Generator(np.random.PCG64(9)).standard_normal((n1, n2, n3)))
    mask = np.asfortranarray(np.random.Generator(np.random.PCG64(10)).random((n1, n2, n3)) < 0.4)
    drv = os.path.join(REF, "traffic_triple_comparison.m")
    it = mlab.Interp(path=[FR])
    unit = it.load(drv, functions_only=True)
    Xhat = out["triple_product"]
    rm, nrm = it.call("evaluate", [Xhat, gt.reshape(-1, order="F")[mask.reshape(-1, order="F")], mask], nargout=2, unit=unit)
    out["evaluate_masked"] = np.array([mlab.scalar(rm), mlab.scalar(nrm)])
    rm, nrm = it.call("evaluate", [Xhat, gt, np.ones(gt.shape, dtype=bool)], nargout=2, unit=unit)
    out["evaluate_all"] = np.array([mlab.scalar(rm), mlab.scalar(nrm)])
    return out


def resolution_summary(it):
    """which file served each function name during the run (MATLAB's resolution order is part of the pin)"""
    seen = {}
    for name, path in it.calls:
        seen.setdefault(name, set()).add(os.path.relpath(path, REF))
    return {k: sorted(v) for k, v in sorted(seen.items())}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ADMM_CASES:
        D, r, o, A0, B0, C0 = case_inputs(name)
        A, B, C, O, eh, text, it = run_admm(D, r, o, A0, B0, C0)
        keep_O = O if O.size <= 60000 else O[:, :, :: max(1, O.shape[2] // 4)]      # big cases: every few slices
        np.savez_compressed(os.path.join(OUT, name + ".npz"), A=A, B=B, C=C, O=keep_O, O_stride=np.array([1 if O.size <= 60000 else max(1, O.shape[2] // 4)]),
                            O_norm=np.array([np.linalg.norm(O.ravel())]), errHist=eh, printed=np.array(text),
                            D_checksum=np.array([D.sum(), np.abs(D).sum()]))
        print(name, "iters", len(eh), "errHist[-1] %.6e" % eh[-1], "| printed", len(text.splitlines()), "lines")
        if name == "tiny_7x6x5_r3":
            print("  function resolution:", resolution_summary(it))
    for name in ALS_CASES:
        X, r, o, A0, B0, C0 = make_golden.case_inputs(name)
        A, B, C, eh, text = run_als(X, r, o, A0, B0, C0)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), A=A, B=B, C=C, errHist=eh, printed=np.array(text),
                            D_checksum=np.array([X.sum(), np.abs(X).sum()]))
        print(name, "iters", len(eh), "errHist[-1] %.6e" % eh[-1])
    np.savez_compressed(os.path.join(OUT, "helpers_7x6x5_r3.npz"), **helper_outputs())
    print("helpers ok")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("make_ref_golden.py needs the reference sources under /root/reference (build container only)")
    main()
