"""Golden vectors of the REFERENCE's own point selection (thirdparty/cvo/src/pcd_generator.cpp +
thirdparty/cvo/thirdparty/PixelSelector2.cpp, compiled where they lie under /root/reference by
`make -C oracle refsel`, run through oracle/ref_select.cpp): for each case the synthetic input is
regenerated from its seed by the tests, the file keeps the reference's outputs — status map (sparse),
selected pixels, positions, features.  Run where /root/reference exists:  python tests/golden/make_refsel_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = {   # name -> (seed, calib name, w, h, high_gradient, num_want, feature_type, gray_mode)
    "tum_default": (21, "tum", 640, 480, False, 3000, 1, 0),
    "eth3d_odd_width": (22, "eth", 739, 458, False, 3000, 1, 0),
    "tum_dense_pot1": (23, "tum", 640, 480, True, 60000, 1, 0),
    "tum_sparse_pot_up": (23, "tum", 640, 480, True, 300, 1, 0),
    "tum_hsv_features_gray14": (24, "tum", 640, 480, False, 3000, 0, 1),
}


def case_input(name):
    seed, cal, w, h, hg, num_want, ft, gm = CASES[name]
    calib = capi.TUM1_CALIB() if cal == "tum" else capi.ETH3D_CALIB()
    scene = synth.make_scene(seed, high_gradient=hg)
    bgr, d = synth.to_numpy(*synth.render(scene, synth.pose(), calib, w, h, noise_seed=seed))
    return calib, bgr, d, num_want, ft, gm


if __name__ == "__main__":
    rs = oracle.load_refsel()
    assert rs is not None, "the reference is not available here"
    out = {}
    for name in CASES:
        calib, bgr, d, num_want, ft, gm = case_input(name)
        r = rs.run(bgr, d, calib, num_want=num_want, feature_type=ft, gray_mode=gm)
        idx = np.flatnonzero(r["map"])
        out[name + "/map_idx"] = idx.astype(np.int32)
        out[name + "/map_val"] = r["map"].reshape(-1)[idx]
        out[name + "/pix"] = r["pix"].astype(np.uint16)          # integer pixel coordinates
        out[name + "/pos"] = r["pos"]
        out[name + "/feat"] = r["feat"]
        out[name + "/input_crc"] = np.array([int(bgr.astype(np.uint64).sum()), int(d.astype(np.uint64).sum())], np.uint64)
        print(name, "points", r["n"], "selected pixels", len(idx))
    path = os.path.join(ROOT, "tests", "golden", "refsel_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
