"""Golden vectors of the REFERENCE's own CVO class (thirdparty/cvo/src/cvo.cpp + LieGroup.cpp + the selection
sources + its vendored nanoflann, compiled where they lie under /root/reference by `make -C oracle refcvo`
against the stand-in headers of oracle/shim/, driven through oracle/ref_cvo.cpp).

For the C1 pair (regenerated from its seed by the tests): cloud sizes after the reference's own set_pcd; at a set
of injected states (R, T, ell) one iteration body of cvo::align — in-cutoff pattern, a_ij, omega, v, step, nnz;
the state after k = 1, 2 free-running iterations; the reference's free-running result; compute_innerproduct at a
fixed transform.  Run where /root/reference exists:  python tests/golden/make_refcvo_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def injected_states(T_final):
    """identity at the four ell of the schedule, then four states near the solution (the inverse of a transform
    close to the final one, perturbed), at the smaller ell"""
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    states = [(I, z, ell) for ell in (0.15, 0.10, 0.06, 0.03)]
    rng = np.random.default_rng(0)
    for k in range(4):
        P = synth.pose(rng.normal(0, 2e-3, 3), rng.normal(0, 2e-3, 3))
        M = np.linalg.inv(np.asarray(T_final, np.float64) @ P)
        states.append((M[:3, :3].astype(np.float32), M[:3, 3].astype(np.float32), (0.10, 0.06, 0.03, 0.03)[k]))
    return states


def keys_of(ij):
    return (ij[:, 0].astype(np.int64) << 16) | ij[:, 1].astype(np.int64)


if __name__ == "__main__":
    cal = capi.TUM1_CALIB()
    a, da, b, db, T_gt = synth.make_pair(1, cal)
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
    for k in (1, 2):   # the first iterations of the free-running loop (before rounding noise can act)
        c = oracle.load_refcvo(cal)
        c.set_pcd(a, da)
        c.set_pcd(b, db)
        c.set_max_iter(k)
        res = c.align()
        R, T, ell, tf = c.get_state()
        out[f"k{k}/R"], out[f"k{k}/T"], out[f"k{k}/ell"] = R, T, np.float32(ell)
        out[f"k{k}/transform"], out[f"k{k}/last_iter_transform"] = res["transform"], res["last_iter_transform"]
        c.close()
    c = oracle.load_refcvo(cal)
    c.set_pcd(a, da)
    c.set_pcd(b, db)
    res = c.align()
    out["free/transform"], out["free/iter"], out["free/nnz"], out["free/ell"] = res["transform"], np.int32(res["iter"]), np.int32(res["A_nonzero"]), np.float32(res["ell"])
    q = c.compute_innerproduct(res["transform"])
    out["free/inn_values"], out["free/inn_nums"], out["free/H"], out["free/inliers"] = q["values"], q["nums"], q["H"], np.int32(q["inliers"])
    out["free/cos_angle"] = np.float32(q["cos_angle"])
    print("free run: iter", res["iter"], "nnz", res["A_nonzero"], "pose error vs ground truth", synth.pose_error(res["transform"], T_gt))
    c.close()
    rc.close()
    path = os.path.join(ROOT, "tests", "golden", "refcvo_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
