"""Generates tests/golden/*.npz from the numpy oracle (oracle/tritd_oracle.py).

The reference is MATLAB and cannot run in the build container (no MATLAB/Octave), and it
ships no fixtures of its own, so these vectors pin the ORACLE, not the MATLAB run: they keep
the oracle from drifting and give the GPU tests committed answers that do not depend on the
numpy/BLAS build of the GPU box.  Inputs are regenerated from seeds (tritd.synth); only
outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import tritd_oracle as orc  # noqa: E402
from tritd import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (config kind, shape, r, iterations, tol)
CASES = {
    "tiny_7x6x5_r3": ("cfg1", (7, 6, 5), 3, 5, 0.0),
    "odd_33x17x9_r2": ("cfg1", (33, 17, 9), 2, 8, 0.0),
    "small_40x36x24_r5": ("cfg1", (40, 36, 24), 5, 20, 0.0),
    "video_48x64x20_r5": ("cfg3", (48, 64, 20), 5, 10, 0.0),
    "traffic_32x32x24_r4": ("cfg2", (32, 32, 24), 4, 10, 0.0),
    "stop_30x30x30_r3": ("cfg1", (30, 30, 30), 3, 100, 2e-2),
}


# ALS (triple_decomp_ALS.m): name -> (config kind, shape, r, iterations, tol); data is the low-rank part plus
# 5 % noise-like outliers so the error history is not trivially zero
ALS_CASES = {
    "als_24x20x16_r3": ("cfg1", (24, 20, 16), 3, 12, 0.0),
    "als_stop_30x28x26_r4": ("cfg1", (30, 28, 26), 4, 60, 1e-3),
}


def case_inputs(name):
    kind, shape, r, iters, tol = (CASES.get(name) or ALS_CASES[name])
    n1, n2, n3, _, k, frac, seed, opts = synth.CONFIGS[kind]
    if k == "lowrank_sparse":
        D = synth.make_lowrank_sparse(*shape, r, frac, seed)
    elif k == "traffic":
        D = synth.make_traffic(*shape, r, frac, seed)
    else:
        D = synth.make_video(*shape, seed)
    A0, B0, C0 = synth.init_factors(*shape, r, 100 + seed)
    o = dict(opts); o["maxIter"] = iters; o["tol"] = tol
    return D, r, o, A0, B0, C0


def main():
    for name in CASES:
        D, r, o, A0, B0, C0 = case_inputs(name)
        A, B, C, O, eh, st = orc.triple_decomp_ADMM(D, r, o, A0, B0, C0, return_state=True)
        L = orc.triple_product(A, B, C)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), A=A, B=B, C=C, O=O, L=L, errHist=eh,
                            errL=st["errL"], errO=st["errO"], D_checksum=np.array([D.sum(), np.abs(D).sum()]))
        print(name, "iters", len(eh), "errHist[-1] %.6e" % eh[-1])
    for name in ALS_CASES:
        X, r, o, A0, B0, C0 = case_inputs(name)
        A, B, C, eh = orc.triple_decomp_ALS(X, r, o, A0, B0, C0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), A=A, B=B, C=C, L=orc.triple_product(A, B, C), errHist=eh,
                            D_checksum=np.array([X.sum(), np.abs(X).sum()]))
        print(name, "iters", len(eh), "errHist[-1] %.6e" % eh[-1])


if __name__ == "__main__":
    main()
