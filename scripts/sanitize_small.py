"""A small tour of every kernel of libcvo_b200.so, sized to run under compute-sanitizer (memcheck / racecheck /
synccheck / initcheck) in seconds: selection, the handle path (cluster of CTAs), the batch path (one CTA per pair,
search + filter + re-search inside 24 iterations), the queries, the batched loop-closure verification and — with
`all` — the dense pair on the cooperative grid.  Results are compared between the handle and the batch path so
that a run which "passes" the sanitizer on garbage is caught as well.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py [small|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cvo_slam_b200 import batch as B, capi, synth

what = sys.argv[1] if len(sys.argv) > 1 else "small"
api = capi.load()
cal = capi.TUM1_CALIB()
scene = synth.make_scene(5)
rng = np.random.default_rng(5)
poses = [synth.pose()] + [synth.pose(rng.normal(0, 6e-3, 3), rng.normal(0, 9e-3, 3)) for _ in range(3)]
frames = [synth.to_numpy(*synth.render(scene, P, cal, 640, 480, noise_seed=50 + k)) for k, P in enumerate(poses)]
bgr = np.stack([f[0] for f in frames])
dep = np.stack([f[1] for f in frames])
pairs = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 0), (1, 3)]

for mode in (0, 1):
    p = api.default_params()
    p.exp_mode = mode
    p.max_iter = 24          # crosses the three length-scale changes (k > 2, 9, 19): search, filter, re-search
    # handle path: a cluster of CTAs on one pair
    h = api.create(cal, p)
    api.set_frame(h, 0, *frames[0])
    api.set_frame(h, 1, *frames[1])
    res, _ = api.align(h)
    ip = api.inner_product(h, 1, res.transform_np(), 0)
    hs = api.hessian(h, 1, res.transform_np(), 0)
    api.destroy(h)
    # batch path: one CTA per pair
    bt = B.Batch(cal, p, max_frames=4, max_pairs=len(pairs), width=640, height=480, api=api)
    bt.set_frames(bgr, dep)
    desc = bt.make_pairs(pairs)
    r = bt.align(desc)
    vals, nums = bt.inner_product(desc, r)
    T = r["transform"].reshape(-1, 4, 4)
    lc = bt.verify_lc(desc, r, T, T, T)
    bt.close()
    assert (r["status"] == 0).all(), r["status"]
    if mode == 0:   # the two paths agree to the bit in the bit-faithful mode
        assert np.array_equal(r["transform"][0].reshape(4, 4), res.transform_np()), "batch and handle path differ"
    print(f"mode {mode}: handle iterations {res.iterations}, batch iterations {r['iterations'].tolist()}, "
          f"inner product {ip[0]:.4f} / {vals[0]:.4f}, lc accept {lc['accept'].tolist()}")

if what == "all":   # the dense pair: cooperative grid, tables in global memory
    a, da, b, db, _ = synth.make_pair(3, cal, high_gradient=True, rot_deg=0.8, trans=(0.015, -0.01, 0.012))
    p = api.default_params()
    p.num_want = 60000
    p.max_iter = 5
    h = api.create(cal, p)
    api.set_frame(h, 0, a, da)
    api.set_frame(h, 1, b, db)
    res, _ = api.align(h)
    print("dense: points", api.slot_size(h, 0), api.slot_size(h, 1), "iterations", res.iterations, "status", res.status)
    api.destroy(h)
print("sanitize tour ok")
