"""Diagnostic: per-iteration trace of the CUDA path beside the oracle on the C1 pair (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from cvo_slam_b200 import capi, synth
from oracle import oracle
from conftest import pose_error

api, orc = capi.load(), oracle.load()
cal = capi.TUM1_CALIB()
a, da, b, db, Tgt = synth.make_pair(1, cal)
res = {}
for name, be in (("cuda", api), ("oracle", orc)):
    h = be.create(cal)
    be.set_frame(h, 0, a, da); be.set_frame(h, 1, b, db)
    r, recs = be.align(h, trace_cap=2000)
    res[name] = (r, recs)
rc, tc = res["cuda"]; ro, to = res["oracle"]
print("iterations", rc.iterations, ro.iterations, "pose diff", pose_error(rc.transform_np(), ro.transform_np()))
for k in range(max(len(tc), len(to))):
    c = tc[k] if k < len(tc) else None
    o = to[k] if k < len(to) else None
    f = lambda r: "%.3e %.3e step %.5f nnz %6d dist %.3e" % (np.linalg.norm(r["omega"]), np.linalg.norm(r["v"]), r["step"], r["nnz"],
                                                    r["step"] * np.sqrt(2 * (r["omega"] ** 2).sum() + (r["v"] ** 2).sum())) if r else "-"
    print(k, "| cuda", f(c), "| oracle", f(o))
# same number of iterations on both sides
for kmax in (10, 20, 30, min(rc.iterations, ro.iterations)):
    T = {}
    for name, be in (("cuda", api), ("oracle", orc)):
        p = be.default_params(); p.max_iter = kmax
        h = be.create(cal, p)
        be.set_frame(h, 0, a, da); be.set_frame(h, 1, b, db)
        r, _ = be.align(h)
        T[name] = (r.transform_np(), r.iterations)
    print("max_iter", kmax, "iterations", T["cuda"][1], T["oracle"][1], "pose diff", pose_error(T["cuda"][0], T["oracle"][0]))
