"""C3 (dense selection stress): kernel-evaluation rate of one large pair, exact and fast mode.
Prints a JSON line per mode with evals/s against the FP32/MUFU roofline of SURVEY §8(d)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_slam_b200 import capi, synth
api = capi.load()
cal = capi.TUM1_CALIB()
a, da, b, db, Tgt = synth.make_pair(3, cal, high_gradient=True, rot_deg=0.8, trans=(0.015, -0.01, 0.012))
SM, CLK = 148, 1.965e9
for mode in (0, 1):
    p = api.default_params(); p.exp_mode = mode; p.num_want = 60000
    h = api.create(cal, p)
    api.set_frame(h, 0, a, da); api.set_frame(h, 1, b, db)
    n = (api.slot_size(h, 0), api.slot_size(h, 1))
    best = None
    for rep in range(4):
        api.set_RT(h, np.eye(3, dtype=np.float32), np.zeros(3, np.float32)); api.set_ell(h, 0.15)
        s0 = api.handle_stats(h)
        t0 = time.perf_counter(); res, _ = api.align(h); dt = time.perf_counter() - t0
        s1 = api.handle_stats(h)
        if best is None or dt < best[0]:
            best = (dt, s1["evals"] - s0["evals"], s1["nnz"] - s0["nnz"], res.iterations, res.status)
    ph = api.phase_cycles(h)
    dt, ev, nnz, it, st = best
    flops = ev * 28 + nnz * 90
    print(json.dumps(dict(workload="C3 dense pair", exp_mode=mode, points=n, iterations=it, status=st, align_ms=dt * 1e3,
                          evals=ev, nnz_sum=nnz, evals_per_s=ev / dt, tflops=flops / dt / 1e12,
                          frac_fp32_peak=flops / dt / (SM * 128 * 2 * CLK),
                          eval_roofline_per_s=min(SM * 128 * 2 * CLK / 28, SM * 16 * CLK / 2),
                          phase_cycles=ph)))
    api.destroy(h)
