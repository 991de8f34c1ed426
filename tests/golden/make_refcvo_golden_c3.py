"""Golden vectors of the REFERENCE's own CVO class on the C3 configuration (BASELINE configs[2]: dense selection, ~18 k
points per cloud): the build of make_refcvo_golden.py (`make -C oracle refcvo`).  The reference's selector is
hard-wired to 3 000 points (pcd_generator.cpp:22), so the dense clouds are handed to its cvo object directly
(positions + features; the dense SELECTION is pinned separately by refsel_golden.npz); one iteration body of
cvo::align at injected states, up to 1.3 M non-zeros.  A pattern of that size is held as its count and the SHA-256
of the sorted (i, j) keys and of the a_ij in that order, beside omega, v, step.
Run where /root/reference exists:  python tests/golden/make_refcvo_golden_c3.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

NUM_WANT = 60000


def c3_pair():
    """the pair of bench.py's c3 leg and of tests/test_gpu_parity.py::test_dense_c3_parity"""
    cal = capi.TUM1_CALIB()
    return (cal,) + tuple(synth.make_pair(3, cal, high_gradient=True, rot_deg=0.8, trans=(0.015, -0.01, 0.012)))


def dense_clouds(api, cal, a, da, b, db):
    """-> handle with both frames set at num_want = 60000, and the clouds (pos, feat) of the two slots"""
    p = api.default_params()
    p.num_want = NUM_WANT
    h = api.create(cal, p)
    api.set_frame(h, 0, a, da)
    api.set_frame(h, 1, b, db)
    return h, api.get_cloud(h, 0), api.get_cloud(h, 1)


def injected_states(T_gt):
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    states = [(I, z, ell) for ell in (0.15, 0.10, 0.06, 0.03)]
    rng = np.random.default_rng(3)
    for k in range(2):
        P = synth.pose(rng.normal(0, 2e-3, 3), rng.normal(0, 2e-3, 3))
        M = np.linalg.inv(np.asarray(T_gt, np.float64) @ P)
        states.append((M[:3, :3].astype(np.float32), M[:3, 3].astype(np.float32), (0.06, 0.03)[k]))
    return states


def pattern_digest(ij, a):
    """(sha256 of the sorted int64 keys i << 16 | j, sha256 of the float32 a_ij in that order) as uint8[32] each"""
    k = (ij[:, 0].astype(np.int64) << 16) | ij[:, 1].astype(np.int64)
    o = np.argsort(k, kind="stable")
    hk = hashlib.sha256(np.ascontiguousarray(k[o]).tobytes()).digest()
    ha = hashlib.sha256(np.ascontiguousarray(a[o].astype(np.float32)).tobytes()).digest()
    return np.frombuffer(hk, np.uint8).copy(), np.frombuffer(ha, np.uint8).copy()


if __name__ == "__main__":
    cal, a, da, b, db, T_gt = c3_pair()
    orc = oracle.load()
    h, (pf, ff), (pm, fm) = dense_clouds(orc, cal, a, da, b, db)
    orc.destroy(h)
    rc = oracle.load_refcvo(cal)
    assert rc is not None, "the reference is not available here"
    rc.set_clouds(pf, ff, pm, fm)
    out = {"sizes": np.array(rc.sizes(), np.int32),
           "input_crc": np.array([int(a.astype(np.uint64).sum()), int(da.astype(np.uint64).sum()),
                                  int(b.astype(np.uint64).sum()), int(db.astype(np.uint64).sum())], np.uint64)}
    states = injected_states(T_gt)
    for s, (R, T, ell) in enumerate(states):
        r = rc.iteration_at(R, T, ell, cap=1 << 23)
        hk, ha = pattern_digest(r["ij"], r["a"])
        out[f"s{s}/R"], out[f"s{s}/T"], out[f"s{s}/ell"] = R, T, np.float32(ell)
        out[f"s{s}/keys_sha256"], out[f"s{s}/a_sha256"] = hk, ha
        out[f"s{s}/omega"], out[f"s{s}/v"] = r["omega"], r["v"]
        out[f"s{s}/step"], out[f"s{s}/nnz"] = np.float32(r["step"]), np.int32(r["nnz"])
        print("state", s, "ell", ell, "nnz", r["nnz"], "omega", r["omega"], "step", r["step"])
    out["n_states"] = np.int32(len(states))
    path = os.path.join(ROOT, "tests", "golden", "refcvo_golden_c3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", flush=True)
    # (the object is left to the end of the process: destroying a reference object that was handed 18 k-point clouds
    # through refcvo_set_clouds faults in the stand-in sparse matrix's teardown — after the results are out)
    os._exit(0)
