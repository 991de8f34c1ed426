"""Times one handle-path alignment (C1 pair) per exp_mode; prints iterations, evals, ms."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_slam_b200 import capi, synth
api = capi.load()
cal = capi.TUM1_CALIB()
a, da, b, db, Tgt = synth.make_pair(1, cal)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for mode in (0, 1):
    p = api.default_params(); p.exp_mode = mode
    h = api.create(cal, p)
    api.set_frame(h, 0, a, da); api.set_frame(h, 1, b, db)
    n = (api.slot_size(h, 0), api.slot_size(h, 1))
    ts = []
    for r in range(reps):
        api.set_RT(h, np.eye(3, dtype=np.float32), np.zeros(3, np.float32)); api.set_ell(h, 0.15)
        s0 = api.handle_stats(h)
        t0 = time.perf_counter(); res, _ = api.align(h); dt = time.perf_counter() - t0
        s1 = api.handle_stats(h)
        ts.append(dt)
    t0 = time.perf_counter()
    for r in range(reps):
        api.set_frame(h, 1, b, db); api.slot_size(h, 1)
    tsel = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for r in range(reps):
        api.inner_product(h, 1, res.transform_np(), 0)
    tq = (time.perf_counter() - t0) / reps
    ph = api.phase_cycles(h); rb = (ph.pop("rebuilds"), ph.pop("filters")); tot = sum(ph.values()); print("   (searches, filters) over all runs of this handle:", rb, "iterations", s1["iterations"])
    print("   phase cycles per iteration:", {k: round(v / max(s1["iterations"], 1)) for k, v in ph.items()}, "total", round(tot / max(s1["iterations"], 1)))
    print(f"exp_mode {mode}: N={n} iterations {res.iterations} evals {s1['evals']-s0['evals']} nnz_sum {s1['nnz']-s0['nnz']} "
          f"align ms min {min(ts)*1e3:.3f} med {sorted(ts)[len(ts)//2]*1e3:.3f} -> {min(ts)*1e6/res.iterations:.1f} us/iter; "
          f"set_frame+sync {tsel*1e3:.3f} ms; inner_product {tq*1e3:.3f} ms")
    api.destroy(h)
