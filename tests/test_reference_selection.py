"""Pins of the point-selection path against the REFERENCE'S OWN CODE.

tests/golden/refsel_golden.npz holds outputs of thirdparty/cvo/src/pcd_generator.cpp +
thirdparty/cvo/thirdparty/PixelSelector2.cpp compiled where they lie (oracle/Makefile `refsel`, stand-in
Eigen / OpenCV headers under oracle/shim/) — made by tests/golden/make_refsel_golden.py.  The inputs are
regenerated from their seeds (and checked against a checksum stored with the vectors).

  * CPU: the oracle's restatement reproduces the reference's status map, pixels, positions and features to the bit;
    where the reference is present the compiled selector is also run live on further frames.
  * GPU: the CUDA selection (through the C ABI) reproduces the same vectors to the bit.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _cases():
    spec = importlib.util.spec_from_file_location("make_refsel_golden", os.path.join(GOLDEN, "make_refsel_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def refsel_golden():
    return dict(np.load(os.path.join(GOLDEN, "refsel_golden.npz")))


def _check_backend(api, g, name, mod):
    calib, bgr, d, num_want, ft, gm = mod.case_input(name)
    crc = g[name + "/input_crc"]
    assert int(bgr.astype(np.uint64).sum()) == int(crc[0]) and int(d.astype(np.uint64).sum()) == int(crc[1]), \
        "the synthetic input of this case is not the one the golden vectors were made from"
    p = api.default_params()
    p.num_want, p.feature_type, p.gray_mode = num_want, ft, gm
    h = api.create(calib, p)
    api.set_frame(h, 0, bgr, d)
    hh, ww = d.shape
    m, info = api.get_selection_debug(h, 0, ww, hh)
    idx = np.flatnonzero(m)
    assert np.array_equal(idx, g[name + "/map_idx"]), name
    assert np.array_equal(m.reshape(-1)[idx], g[name + "/map_val"]), name
    pix = api.get_selected_points(h, 0)
    assert np.array_equal(pix, g[name + "/pix"].astype(np.float32)), name
    pos, feat = api.get_cloud(h, 0)
    assert np.array_equal(pos.view(np.uint32), g[name + "/pos"].view(np.uint32)), name
    assert np.array_equal(feat.view(np.uint32), g[name + "/feat"].view(np.uint32)), name
    api.destroy(h)
    return len(pos), info


@pytest.mark.parametrize("name", ["tum_default", "eth3d_odd_width", "tum_dense_pot1", "tum_sparse_pot_up",
                                  "tum_hsv_features_gray14"])
def test_oracle_matches_reference_selection_golden(oracle_api, refsel_golden, name):
    n, info = _check_backend(oracle_api, refsel_golden, name, _cases())
    if name == "tum_dense_pot1":
        assert info["passes"] == 2 and info["pot"] == 1     # makeMaps recursion, quotia > 1.25
    if name == "tum_sparse_pot_up":
        assert info["passes"] == 2 and info["pot"] > 3      # makeMaps recursion, quotia < 0.25


def test_reference_selector_live_vs_oracle(oracle_api, tum_calib, pair_c1):
    """Where /root/reference is present (this container; the built library travels to the GPU box): the
    reference's compiled selector, run live on the C1 frames and on a flat image, against the oracle."""
    from oracle import oracle
    rs = oracle.load_refsel()
    if rs is None:
        pytest.skip("oracle/_ref/libref_select.so not built (no /root/reference here)")
    flat = (np.full((480, 640, 3), 90, np.uint8), np.full((480, 640), 5000, np.uint16))
    for bgr, d in ((pair_c1[0], pair_c1[1]), (pair_c1[2], pair_c1[3]), flat):
        r = rs.run(bgr, d, tum_calib)
        h = oracle_api.create(tum_calib)
        oracle_api.set_frame(h, 0, bgr, d)
        m, _ = oracle_api.get_selection_debug(h, 0, 640, 480)
        assert np.array_equal(m, r["map"])
        if r["n"]:
            pos, feat = oracle_api.get_cloud(h, 0)
            assert np.array_equal(oracle_api.get_selected_points(h, 0), r["pix"])
            assert np.array_equal(pos.view(np.uint32), r["pos"].view(np.uint32))
            assert np.array_equal(feat.view(np.uint32), r["feat"].view(np.uint32))
        else:
            assert oracle_api.slot_size(h, 0) == 0
        oracle_api.destroy(h)
    # the stand-in cvtColor of the shim is the arithmetic pinned against cv2 (tests/test_oracle_pins.py)
    cv2 = pytest.importorskip("cv2")
    r = rs.run(pair_c1[0], pair_c1[1], tum_calib)
    assert np.array_equal(r["gray"], cv2.cvtColor(pair_c1[0], cv2.COLOR_RGB2GRAY))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tum_default", "eth3d_odd_width", "tum_dense_pot1", "tum_sparse_pot_up",
                                  "tum_hsv_features_gray14"])
def test_cuda_matches_reference_selection_golden(cuda_api, refsel_golden, name):
    _check_backend(cuda_api, refsel_golden, name, _cases())
