"""Generates tests/golden/final_rre.json: the outcome of the reference's FULL run on the BASELINE configs from the CPU oracle
(multi-threaded port oracle/tritd_oracle_mt.py): [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts) with the reference's
own options (tol 1e-5, maxIter 100), then the drivers' RRE  ||triple_product(A,B,C) - L0||_F / ||L0||_F
(traffic_triple_comparison.m:194-199; L0 = the low-rank part the synthetic D was built from), the executed iteration count
and the last errHist value.  bench.py prints the same three numbers from the GPU run next to these (`final_rre`) -- north_star:
"identical iteration count to convergence, and the final RRE reported" -- without running the CPU oracle on the GPU box.
    python tests/golden/make_final_rre.py [cfg ...]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"))
import tritd_oracle as orc  # noqa: E402
import tritd_oracle_mt as mt  # noqa: E402
from tritd import synth  # noqa: E402

OUT = os.path.join(HERE, "final_rre.json")
res = json.load(open(OUT)) if os.path.exists(OUT) else {}
for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]):
    w = synth.make_config(name, with_truth=True)
    o = dict(w["opts"], disp=0)
    t = time.time()
    A, B, C, O, eh = mt.triple_decomp_ADMM(w["D"], w["r"], o, w["A0"], w["B0"], w["C0"])
    res[name] = {"shape": list(w["shape"]), "r": w["r"], "RRE": synth.rre(orc.triple_product(A, B, C), w["L0"]),
                 "iterations": int(len(eh)), "final_errHist": float(eh[-1]), "tol": o["tol"], "maxIter": int(o["maxIter"])}
    print(name, w["shape"], "%.1f s" % (time.time() - t), res[name], flush=True)
    json.dump(res, open(OUT, "w"), indent=1)
