"""Golden vectors of the REFERENCE's own cvo::compute_innerproduct_lc (cvo.cpp:505-561; the build of
make_refcvo_golden.py, `make -C oracle refcvo`) on the C1 pair: the loop-closure verification record — six inner
products with their pair counts, the eigenvalue-shifted Hessian, the two inlier counts, cos_angle — for three sets of
candidate transforms at the three length scales the schedule can leave an object at (0.10, 0.06, 0.03).
Run where /root/reference exists:  python tests/golden/make_refcvo_golden_lc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def candidate_sets(T_gt):
    """(ell, prior, lc_prior, lc_prior_2, lc) x 3: lc = the ground truth under a small perturbation (what an alignment
    returns), the priors = coarser estimates (motion model / PnP-RANSAC stand-ins)"""
    rng = np.random.default_rng(12)
    sets = []
    for ell, s_lc, s_pr in ((0.10, 3e-3, 2e-2), (0.06, 1e-3, 1e-2), (0.03, 3e-4, 5e-3)):
        def near(scale):
            return (np.asarray(T_gt, np.float64) @ synth.pose(rng.normal(0, scale, 3), rng.normal(0, scale, 3))).astype(np.float32)
        sets.append((ell, near(s_pr), near(s_pr), near(s_pr / 2), near(s_lc)))
    return sets


if __name__ == "__main__":
    cal = capi.TUM1_CALIB()
    a, da, b, db, T_gt = synth.make_pair(1, cal)
    rc = oracle.load_refcvo(cal)
    assert rc is not None, "the reference is not available here"
    rc.set_pcd(a, da)
    rc.set_pcd(b, db)
    out = {"input_crc": np.array([int(a.astype(np.uint64).sum()), int(da.astype(np.uint64).sum()),
                                  int(b.astype(np.uint64).sum()), int(db.astype(np.uint64).sum())], np.uint64)}
    sets = candidate_sets(T_gt)
    for s, (ell, prior, lc_prior, lc_prior_2, lc) in enumerate(sets):
        rc.set_state(np.eye(3, dtype=np.float32), np.zeros(3, np.float32), ell)
        r = rc.compute_innerproduct_lc(prior, lc_prior, lc_prior_2, lc)
        out[f"c{s}/ell"] = np.float32(ell)
        out[f"c{s}/prior"], out[f"c{s}/lc_prior"], out[f"c{s}/lc_prior_2"], out[f"c{s}/lc"] = prior, lc_prior, lc_prior_2, lc
        out[f"c{s}/values"], out[f"c{s}/nums"], out[f"c{s}/H"] = r["values"], r["nums"], r["H"]
        out[f"c{s}/inliers_svd"], out[f"c{s}/inliers_pnpransac"] = np.int32(r["inliers_svd"]), np.int32(r["inliers_pnpransac"])
        out[f"c{s}/cos_angle"] = np.float32(r["cos_angle"])
        print("set", s, "ell", ell, "values", r["values"], "nums", r["nums"], "inliers", r["inliers_svd"], r["inliers_pnpransac"],
              "cos", r["cos_angle"])
    out["n_sets"] = np.int32(len(sets))
    rc.close()
    path = os.path.join(ROOT, "tests", "golden", "refcvo_golden_lc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
