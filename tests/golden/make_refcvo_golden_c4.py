"""Golden vectors of the REFERENCE's own CVO class on the C4 configuration (BASELINE configs[3]: ETH3D-shaped 739x458
pair, 8 deg / 0.15 m initial motion, wide cutoff): the same build of thirdparty/cvo/src/cvo.cpp + LieGroup.cpp + the
selection sources + nanoflann as make_refcvo_golden.py (`make -C oracle refcvo`), driven through oracle/ref_cvo.cpp.

Held: cloud sizes after the reference's own set_pcd (odd image width: the stride quirk of pcd_generator.cpp:103); one
iteration body of cvo::align at injected states — identity at ell 0.25 / 0.15 / 0.10 / 0.06 (far from the solution:
the sparse, wide-cutoff regime) and four states near the ground truth at ell 0.25 / 0.10 / 0.06 / 0.03 (dense regime) —
with the in-cutoff pattern, every a_ij, omega, v, step, nnz; the state after k = 1, 2, 3 free-running iterations from
ell_init = 0.25; inner products and the Hessian at the ground-truth transform.  The free-running result itself is not
held: from this start the schedule leaves the basin, so it is defined only up to chaos.
Run where /root/reference exists:  python tests/golden/make_refcvo_golden_c4.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

ELL_INIT = 0.25


def c4_pair():
    """the pair of bench.py's c4 leg and of tests/test_gpu_parity.py::test_align_c4_specified_large_motion"""
    eth = capi.ETH3D_CALIB()
    t = np.array([0.10, -0.05, 0.10])
    t = t / np.linalg.norm(t) * 0.15
    return (eth,) + tuple(synth.make_pair(4, eth, w=739, h=458, rot_deg=8.0, trans=tuple(t)))


def injected_states(T_gt):
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    states = [(I, z, ell) for ell in (0.25, 0.15, 0.10, 0.06)]
    rng = np.random.default_rng(4)
    for k in range(4):
        P = synth.pose(rng.normal(0, 2e-3, 3), rng.normal(0, 2e-3, 3))
        M = np.linalg.inv(np.asarray(T_gt, np.float64) @ P)
        states.append((M[:3, :3].astype(np.float32), M[:3, 3].astype(np.float32), (0.25, 0.10, 0.06, 0.03)[k]))
    return states


def keys_of(ij):
    return (ij[:, 0].astype(np.int64) << 16) | ij[:, 1].astype(np.int64)


if __name__ == "__main__":
    cal, a, da, b, db, T_gt = c4_pair()
    rc = oracle.load_refcvo(cal)
    assert rc is not None, "the reference is not available here"
    rc.set_pcd(a, da)
    rc.set_pcd(b, db)
    out = {"sizes": np.array(rc.sizes(), np.int32),
           "input_crc": np.array([int(a.astype(np.uint64).sum()), int(da.astype(np.uint64).sum()),
                                  int(b.astype(np.uint64).sum()), int(db.astype(np.uint64).sum())], np.uint64)}
    states = injected_states(T_gt)
    for s, (R, T, ell) in enumerate(states):
        r = rc.iteration_at(R, T, ell)
        k = keys_of(r["ij"])
        o = np.argsort(k)
        out[f"s{s}/R"], out[f"s{s}/T"], out[f"s{s}/ell"] = R, T, np.float32(ell)
        out[f"s{s}/keys"] = k[o]
        out[f"s{s}/a"] = r["a"][o]
        out[f"s{s}/omega"], out[f"s{s}/v"] = r["omega"], r["v"]
        out[f"s{s}/step"], out[f"s{s}/nnz"] = np.float32(r["step"]), np.int32(r["nnz"])
        print("state", s, "ell", ell, "nnz", r["nnz"], "omega", r["omega"], "step", r["step"])
    out["n_states"] = np.int32(len(states))
    for k in (1, 2, 3):
        c = oracle.load_refcvo(cal)
        c.set_pcd(a, da)
        c.set_pcd(b, db)
        c.set_state(np.eye(3, dtype=np.float32), np.zeros(3, np.float32), ELL_INIT)   # a fresh object whose ell_init is 0.25
        c.set_max_iter(k)
        res = c.align()
        R, T, ell, tf = c.get_state()
        out[f"k{k}/R"], out[f"k{k}/T"], out[f"k{k}/ell"] = R, T, np.float32(ell)
        out[f"k{k}/transform"], out[f"k{k}/last_iter_transform"] = res["transform"], res["last_iter_transform"]
        print("after", k, "iteration(s): ell", ell, "T", T)
        c.close()
    # queries at the ground-truth transform, at the last length scale of the schedule
    rc.set_state(np.eye(3, dtype=np.float32), np.zeros(3, np.float32), 0.03)
    Tq = np.asarray(T_gt, np.float32)
    q = rc.compute_innerproduct(Tq)
    out["gt/transform"] = Tq
    out["gt/inn_values"], out["gt/inn_nums"], out["gt/H"], out["gt/inliers"] = q["values"], q["nums"], q["H"], np.int32(q["inliers"])
    print("queries at the ground truth:", q["values"], q["nums"], "inliers", q["inliers"])
    rc.close()
    path = os.path.join(ROOT, "tests", "golden", "refcvo_golden_c4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
