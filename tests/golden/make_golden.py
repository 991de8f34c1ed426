"""Generates tests/golden/pair_640x480_seed11.npz from the CPU oracle.

The reference ships no golden vectors and cannot be built here, so these are
produced by the oracle whose radius search is the reference's own nanoflann
(oracle/_ref/libcvo_oracle_kd.so).  They serve two purposes: (1) pin the oracle against
regressions on any machine (tests/test_oracle_pins.py), (2) let the CUDA path be checked
against fixed vectors (tests/test_gpu_parity.py).

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cvo_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

W, H, SEED = 640, 480, 11


def calib_small():
    return capi.TUM1_CALIB()


def main():
    o = oracle.load(kd=True)
    assert o.has_nanoflann
    cal = calib_small()
    # sensor noise sigma 1 (instead of 2) keeps the compressed fixture near 1.5 MB
    bgr_a, d_a, bgr_b, d_b, T_gt = synth.make_pair(SEED, cal, w=W, h=H, rot_deg=0.8,
                                                   trans=(0.015, -0.01, 0.012), noise_sigma=1.0)
    out = dict(bgr_a=bgr_a, depth_a=d_a, bgr_b=bgr_b, depth_b=d_b, T_gt=T_gt,
               calib=np.array([cal.scaling_factor, cal.fx, cal.fy, cal.cx, cal.cy], np.float32))
    h = o.create(cal)
    o.set_frame(h, 0, bgr_a, d_a)
    o.set_frame(h, 1, bgr_b, d_b)
    for name, slot in (("a", 0), ("b", 1)):
        m, info = o.get_selection_debug(h, slot, W, H)
        out[f"map_idx_{name}"] = np.flatnonzero(m).astype(np.int32)
        out[f"map_val_{name}"] = m.reshape(-1)[np.flatnonzero(m)].astype(np.uint8)
        out[f"sel_info_{name}"] = np.array([info[k] for k in ("n2", "n3", "n4", "pot", "passes")], np.int32)
        pos, feat = o.get_cloud(h, slot)
        out[f"pos_{name}"] = pos
        out[f"feat_{name}"] = feat
        out[f"pix_{name}"] = o.get_selected_points(h, slot)
    st = o.stages(bgr_a)
    out["gray_a"] = st["gray"]
    out["ths_smoothed_a"] = st["ths_smoothed"]

    # injected-state iterations (per-iteration parity, SURVEY §8d "Parity gates")
    states = []
    I = np.eye(3, dtype=np.float32)
    z = np.zeros(3, np.float32)
    for k, (R, T, ell) in enumerate([(I, z, 0.15), (I, z, 0.10), (I, z, 0.03)]):
        rec = o.iteration_at(h, R, T, ell)
        ij, a, n = o.last_pattern(h, 400000)
        assert n == len(a)
        out[f"it{k}_ell"] = np.float32(ell)
        out[f"it{k}_omega"] = rec["omega"]
        out[f"it{k}_v"] = rec["v"]
        out[f"it{k}_BCDE"] = np.array([rec["B"], rec["C"], rec["D"], rec["E"]])
        out[f"it{k}_step"] = np.float32(rec["step"])
        out[f"it{k}_nnz"] = np.int32(rec["nnz"])
        if k > 0:   # the ell = 0.15 pattern is the largest; its count and sums are kept
            out[f"it{k}_ij"] = ij.astype(np.uint16)
            out[f"it{k}_a"] = a
        states.append(rec)

    res, recs = o.align(h, trace_cap=2000)
    out["align_transform"] = res.transform_np()
    out["align_R"] = res.R_np()
    out["align_T"] = res.T_np()
    out["align_scalars"] = np.array([res.iterations, res.iter, res.A_nonzero], np.int32)
    out["align_ell"] = np.float32(res.ell)
    out["trace_omega"] = np.array([r["omega"] for r in recs], np.float32)
    out["trace_v"] = np.array([r["v"] for r in recs], np.float32)
    out["trace_step"] = np.array([r["step"] for r in recs], np.float32)
    out["trace_nnz"] = np.array([r["nnz"] for r in recs], np.int32)
    out["trace_ell"] = np.array([r["ell"] for r in recs], np.float32)

    T = res.transform_np()
    ips = [o.inner_product(h, 1, None, 0), o.inner_product(h, 1, T, 0),
           o.inner_product(h, 0, None, 0), o.inner_product(h, 1, None, 1)]
    out["inner_values"] = np.array([v for v, _ in ips], np.float32)
    out["inner_nums"] = np.array([n for _, n in ips], np.int32)
    Hm, inl = o.hessian(h, 1, T, 0)
    out["hessian"] = Hm
    out["hessian_inliers"] = np.int32(inl)
    path = os.path.join(ROOT, "tests", "golden", f"pair_{W}x{H}_seed{SEED}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;",
          "N =", len(out["pos_a"]), len(out["pos_b"]), "iterations", res.iterations)


if __name__ == "__main__":
    main()
