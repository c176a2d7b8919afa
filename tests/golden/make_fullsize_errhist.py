"""Generates tests/golden/fullsize_errhist.json: the first errHist values of the BASELINE configs at FULL size from
the CPU oracle (multi-threaded port oracle/tritd_oracle_mt.py, itself checked against the numpy oracle and, through
it, against the reference's own .m source).  errHist(k) = ||resL||/||D|| + ||resO||/||D|| (:59) depends on every
array of the iteration (A, B, C, L, O, E, Y_L), so a GPU run that reproduces these values at full size has made the
same iterates -- without having to run the CPU oracle on the GPU box.  Used by tests/test_gpu_fullsize.py and by
bench.py (`parity_vs_fixture`, also at N > 1 ranks).   python tests/golden/make_fullsize_errhist.py [cfg ...]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import tritd_oracle_mt as mt  # noqa: E402
from tritd import synth  # noqa: E402

ITERS = {"cfg1": 10, "cfg2": 5, "cfg3": 5, "cfg4": 3}
OUT = os.path.join(HERE, "fullsize_errhist.json")

res = json.load(open(OUT)) if os.path.exists(OUT) else {}
for name in (sys.argv[1:] or list(ITERS)):
    w = synth.make_config(name)
    o = dict(w["opts"], maxIter=ITERS[name], tol=0.0, disp=0)
    t = time.time()
    A, B, C, O, eh = mt.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    res[name] = {"shape": list(w["shape"]), "r": w["r"], "errHist": [float(x) for x in eh],
                 "opts": {k: o[k] for k in ("mu", "rho", "lambda", "lambda2")}}
    print(name, w["shape"], "%.1f s" % (time.time() - t), eh)
    json.dump(res, open(OUT, "w"), indent=1)
