"""Writes the inputs of tools/reference_dump.m (tests/golden/dump_in_<case>.mat, MATLAB v5 format): the seeded
tensors and initial factors of the pinned cases, so that a machine with MATLAB / Octave reproduces exactly the inputs
the oracle and the GPU tests use.  python tools/reference_dump_inputs.py [case ...]"""
import os
import sys

import numpy as np
from scipy.io import savemat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "triple-tensor-decomposition-with-admm_b200", os.path.join("tests", "golden")):
    sys.path.insert(0, os.path.join(ROOT, p))
import make_ref_golden as mrg  # noqa: E402

CASES = {"cfg1": "cfg1_50x50x50_r5_full", "tiny": "tiny_7x6x5_r3", "stop": "stop_30x30x30_r3"}

for short in (sys.argv[1:] or list(CASES)):
    D, r, o, A0, B0, C0 = mrg.case_inputs(CASES[short])
    opts = {k: float(o[k]) for k in ("mu", "rho", "lambda", "lambda2", "maxIter", "tol")}
    opts["disp"] = 0.0
    out = os.path.join(ROOT, "tests", "golden", f"dump_in_{short}.mat")
    savemat(out, dict(D=D, r=float(r), opts=opts, A0=A0, B0=B0, C0=C0), format="5", do_compression=True)
    print("wrote", out, D.shape, "r =", r)
