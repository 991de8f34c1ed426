"""Parity of the CUDA path (libcvo_b200.so, through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star / SURVEY §8d "Parity gates"):
  * point selection, pixel coordinates, positions, features: bit-exact
  * default (exact) mode, injected (R, T, ell): identical in-cutoff pattern, a_ij, omega, v to the
    bit; B..E to 1e-12 relative (double sums in a different order); step to 1e-6
  * default mode, free running: same iteration count as the oracle, final pose within
    1e-4 rad / 1e-4 m (observed ~1e-7)
  * fast mode (exp_mode = 1, MUFU ex2): pattern identical except ties |a - sp|/sp < 1e-5,
    a_ij to 3e-6, omega / v to 1e-5, B..E and step to 1e-4; the free-running trajectory
    decorrelates from the oracle's in the chaotic tail, so its final pose is
    only required to sit in the same basin
  * inner products: 1e-4 relative; Hessian: 1e-4 of its largest entry
"""
import numpy as np
import pytest

from conftest import pose_error

pytestmark = pytest.mark.gpu

POSE_TOL_RAD = 1e-4
POSE_TOL_M = 1e-4
INNER_RTOL = 1e-4


def _calib_of(g):
    from cvo_slam_b200 import capi
    return capi.Calib(*[float(x) for x in g["calib"]])


def _both(cuda_api, oracle_api, calib, params=None):
    return cuda_api.create(calib, params), oracle_api.create(calib, params)


def test_library_and_random_pattern(cuda_api, oracle_api):
    n = 640 * 480
    assert np.array_equal(cuda_api.random_pattern(n), oracle_api.random_pattern(n))


def _check_selection(cuda_api, oracle_api, calib, bgr, depth, params=None):
    hc, ho = _both(cuda_api, oracle_api, calib, params)
    h, w = depth.shape
    cuda_api.set_frame(hc, 0, bgr, depth)
    oracle_api.set_frame(ho, 0, bgr, depth)
    mo, io = oracle_api.get_selection_debug(ho, 0, w, h)
    mc, ic = cuda_api.get_selection_debug(hc, 0, w, h)
    assert ic == io
    assert np.array_equal(mc, mo), f"{int((mc != mo).sum())} status-map pixels differ"
    assert cuda_api.slot_size(hc, 0) == oracle_api.slot_size(ho, 0)
    assert np.array_equal(cuda_api.get_selected_points(hc, 0), oracle_api.get_selected_points(ho, 0))
    pc, fc = cuda_api.get_cloud(hc, 0)
    po, fo = oracle_api.get_cloud(ho, 0)
    assert np.array_equal(pc.view(np.uint32), po.view(np.uint32))
    assert np.array_equal(fc.view(np.uint32), fo.view(np.uint32))
    n = len(pc)
    cuda_api.destroy(hc)
    oracle_api.destroy(ho)
    return n, io


def test_selection_bit_exact_c1(cuda_api, oracle_api, tum_calib, pair_c1):
    bgr_a, d_a, bgr_b, d_b, _ = pair_c1
    for bgr, d in ((bgr_a, d_a), (bgr_b, d_b)):
        n, info = _check_selection(cuda_api, oracle_api, tum_calib, bgr, d)
        assert 2000 < n < 3400


def test_selection_bit_exact_golden(cuda_api, golden_small):
    g = golden_small
    h = cuda_api.create(_calib_of(g))
    cuda_api.set_frame(h, 0, g["bgr_a"], g["depth_a"])
    m, info = cuda_api.get_selection_debug(h, 0, 640, 480)
    assert np.array_equal(np.flatnonzero(m), g["map_idx_a"])
    assert np.array_equal(m.reshape(-1)[g["map_idx_a"]], g["map_val_a"])
    assert [info[k] for k in ("n2", "n3", "n4", "pot", "passes")] == g["sel_info_a"].tolist()
    pos, feat = cuda_api.get_cloud(h, 0)
    assert np.array_equal(pos, g["pos_a"]) and np.array_equal(feat, g["feat_a"])
    assert np.array_equal(cuda_api.get_selected_points(h, 0), g["pix_a"])
    cuda_api.destroy(h)


def test_selection_eth3d_shape_odd_width(cuda_api, oracle_api):
    """739 x 458: odd-width pyramid stride quirk and the out-of-range threshold rows."""
    from cvo_slam_b200 import capi, synth
    cal = capi.ETH3D_CALIB()
    scene = synth.make_scene(4)
    bgr, d = synth.to_numpy(*synth.render(scene, synth.pose(), cal, 739, 458, noise_seed=4))
    n, info = _check_selection(cuda_api, oracle_api, cal, bgr, d)
    assert n > 1500


def test_selection_dense_and_sparse_recursion(cuda_api, oracle_api, tum_calib):
    """num_want = 60000 drives the selector to pot = 1 (quotia > 1.25); num_want = 300 to a larger
    pot (quotia < 0.25) — both branches of makeMaps' recursion (PixelSelector2.cpp:193-223)."""
    from cvo_slam_b200 import synth
    scene = synth.make_scene(3, high_gradient=True)
    bgr, d = synth.to_numpy(*synth.render(scene, synth.pose(), tum_calib, 640, 480, noise_seed=3))
    p = cuda_api.default_params()
    p.num_want = 60000
    n, info = _check_selection(cuda_api, oracle_api, tum_calib, bgr, d, p)
    assert info["passes"] == 2 and info["pot"] == 1 and n > 8000
    p.num_want = 300
    n, info = _check_selection(cuda_api, oracle_api, tum_calib, bgr, d, p)
    assert info["passes"] == 2 and info["pot"] > 3


def test_selection_feature_type0_and_gray14(cuda_api, oracle_api, tum_calib, pair_c1):
    bgr_a, d_a = pair_c1[0], pair_c1[1]
    p = cuda_api.default_params()
    p.feature_type = 0
    _check_selection(cuda_api, oracle_api, tum_calib, bgr_a, d_a, p)
    p = cuda_api.default_params()
    p.gray_mode = 1
    _check_selection(cuda_api, oracle_api, tum_calib, bgr_a, d_a, p)


def test_selection_flat_image_selects_nothing(cuda_api, oracle_api, tum_calib):
    bgr = np.full((480, 640, 3), 90, np.uint8)
    d = np.full((480, 640), 5000, np.uint16)
    n, info = _check_selection(cuda_api, oracle_api, tum_calib, bgr, d)
    assert n == 0


def _pattern_keys(ij):
    return ij[:, 0].astype(np.int64) * (1 << 20) + ij[:, 1].astype(np.int64)


def _check_iteration(cuda_api, oracle_api, hc, ho, R, T, ell, sp=8e-3, exact=True):
    rc = cuda_api.iteration_at(hc, R, T, ell)
    ro = oracle_api.iteration_at(ho, R, T, ell)
    ijc, ac, nc = cuda_api.last_pattern(hc, 1 << 20)
    ijo, ao, no = oracle_api.last_pattern(ho, 1 << 20)
    assert nc == len(ac) == rc["nnz"] and no == ro["nnz"]
    kc, ko = _pattern_keys(ijc), _pattern_keys(ijo)
    oc, oo = np.argsort(kc), np.argsort(ko)
    kc, ac, ko, ao = kc[oc], ac[oc], ko[oo], ao[oo]
    if exact:
        assert np.array_equal(kc, ko), "in-cutoff pattern differs"
        assert np.array_equal(ac.view(np.uint32), ao.view(np.uint32)), \
            f"{int((ac != ao).sum())} of {len(ao)} kernel values differ in the last bits"
        assert np.array_equal(rc["omega"].view(np.uint32), ro["omega"].view(np.uint32))
        assert np.array_equal(rc["v"].view(np.uint32), ro["v"].view(np.uint32))
        for key in "BCDE":
            assert rc[key] == pytest.approx(ro[key], rel=1e-12, abs=1e-300), key
        assert rc["step"] == pytest.approx(ro["step"], rel=1e-6)
        return 0, no
    only_c = np.setdiff1d(kc, ko, assume_unique=True)
    only_o = np.setdiff1d(ko, kc, assume_unique=True)
    # documented cutoff ties: MUFU ex2 vs double exp can flip `a > sp_thres` within 1e-5 relative
    for k in only_c:
        assert abs(ac[np.searchsorted(kc, k)] - sp) / sp < 1e-5
    for k in only_o:
        assert abs(ao[np.searchsorted(ko, k)] - sp) / sp < 1e-5
    ties = len(only_c) + len(only_o)
    assert ties <= max(3, no // 2000), (ties, no)
    common = np.intersect1d(kc, ko, assume_unique=True)
    a_c = ac[np.searchsorted(kc, common)]
    a_o = ao[np.searchsorted(ko, common)]
    assert np.allclose(a_c, a_o, rtol=3e-6, atol=0)
    scale_w = max(np.abs(ro["omega"]).max(), 1e-6)
    scale_v = max(np.abs(ro["v"]).max(), 1e-6)
    assert np.allclose(rc["omega"], ro["omega"], rtol=1e-5, atol=2e-5 * scale_w)
    assert np.allclose(rc["v"], ro["v"], rtol=1e-5, atol=2e-5 * scale_v)
    for key in "BCDE":
        assert rc[key] == pytest.approx(ro[key], rel=1e-4, abs=1e-4 * max(abs(ro["B"]), 1e-9)), key
    assert rc["step"] == pytest.approx(ro["step"], rel=1e-4)
    return ties, no


def _injected_states(g):
    from cvo_slam_b200 import synth
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    states = [(I, z, ell) for ell in (0.15, 0.10, 0.06, 0.03)]
    Tf = g["align_transform"].astype(np.float64)
    rng = np.random.default_rng(0)
    for k in range(4):   # near the solution: inverse of the oracle's final transform, perturbed
        P = synth.pose(rng.normal(0, 2e-3, 3), rng.normal(0, 2e-3, 3))
        M = np.linalg.inv(Tf @ P)
        states.append((M[:3, :3].astype(np.float32), M[:3, 3].astype(np.float32), (0.10, 0.06, 0.03, 0.03)[k]))
    return states


@pytest.mark.parametrize("exp_mode", [0, 1])
def test_iteration_parity_golden(cuda_api, oracle_api, golden_small, exp_mode):
    g = golden_small
    p = cuda_api.default_params()
    p.exp_mode = exp_mode
    hc, ho = cuda_api.create(_calib_of(g), p), oracle_api.create(_calib_of(g))
    for api, h in ((cuda_api, hc), (oracle_api, ho)):
        api.set_cloud(h, 0, g["pos_a"], g["feat_a"])
        api.set_cloud(h, 1, g["pos_b"], g["feat_b"])
    total_ties = 0
    for R, T, ell in _injected_states(g):
        ties, n = _check_iteration(cuda_api, oracle_api, hc, ho, R, T, ell, exact=(exp_mode == 0))
        total_ties += ties
    # golden vectors (made with the reference's nanoflann) at the identity state
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    rec = cuda_api.iteration_at(hc, I, z, 0.15)
    if exp_mode == 0:
        assert rec["nnz"] == int(g["it0_nnz"])
        assert np.array_equal(rec["omega"], g["it0_omega"]) and np.array_equal(rec["v"], g["it0_v"])
    else:
        assert abs(rec["nnz"] - int(g["it0_nnz"])) <= 2
        assert np.allclose(rec["omega"], g["it0_omega"], rtol=1e-5, atol=1e-5 * np.abs(g["it0_omega"]).max())
    print("exp_mode", exp_mode, "cutoff ties over 8 injected states:", total_ties)
    cuda_api.destroy(hc)
    oracle_api.destroy(ho)


def test_align_parity_golden(cuda_api, golden_small):
    g = golden_small
    h = cuda_api.create(_calib_of(g))
    cuda_api.set_frame(h, 0, g["bgr_a"], g["depth_a"])
    cuda_api.set_frame(h, 1, g["bgr_b"], g["depth_b"])
    res, recs = cuda_api.align(h, trace_cap=2000)
    ang, dist = pose_error(res.transform_np(), g["align_transform"])
    print("golden: iterations gpu/oracle", res.iterations, int(g["align_scalars"][0]), "pose diff", ang, dist)
    assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
    assert res.status == 0
    assert res.ell == pytest.approx(float(g["align_ell"]))
    assert [res.iterations, res.iter, res.A_nonzero] == g["align_scalars"].tolist()
    n = len(recs)
    assert np.array_equal([r["ell"] for r in recs], g["trace_ell"][:n])
    # the whole trajectory follows the golden trace
    assert np.array_equal([r["nnz"] for r in recs], g["trace_nnz"][:n])
    assert np.allclose(np.array([r["omega"] for r in recs]), g["trace_omega"][:n], rtol=1e-5, atol=1e-9)
    assert np.allclose(np.array([r["step"] for r in recs]), g["trace_step"][:n], rtol=1e-5)
    # queries at the final state
    T = res.transform_np()
    vals = [cuda_api.inner_product(h, 1, None, 0), cuda_api.inner_product(h, 1, T, 0),
            cuda_api.inner_product(h, 0, None, 0), cuda_api.inner_product(h, 1, None, 1)]
    for (v, n_), gv, gn, name in zip(vals, g["inner_values"], g["inner_nums"], ("pre", "post", "fixed", "moving")):
        assert v == pytest.approx(float(gv), rel=INNER_RTOL), name
        assert n_ == int(gn), name
    cuda_api.destroy(h)


def test_queries_parity_same_pose(cuda_api, oracle_api, golden_small):
    """inner products and Hessian at exactly the same transform and ell on both sides."""
    g = golden_small
    hc, ho = _both(cuda_api, oracle_api, _calib_of(g))
    for api, h in ((cuda_api, hc), (oracle_api, ho)):
        api.set_cloud(h, 0, g["pos_a"], g["feat_a"])
        api.set_cloud(h, 1, g["pos_b"], g["feat_b"])
    T = g["align_transform"]
    for ell in (0.15, 0.03):
        cuda_api.set_ell(hc, ell)
        oracle_api.set_ell(ho, ell)
        for sa, Ta, sb in ((1, None, 0), (1, T, 0), (0, None, 0), (1, None, 1)):
            vc, nc = cuda_api.inner_product(hc, sa, Ta, sb)
            vo, no = oracle_api.inner_product(ho, sa, Ta, sb)
            assert nc == no
            assert vc == pytest.approx(vo, rel=INNER_RTOL)
        Hc, ic = cuda_api.hessian(hc, 1, T, 0)
        Ho, io = oracle_api.hessian(ho, 1, T, 0)
        assert ic == io
        assert np.allclose(Hc, Ho, rtol=0, atol=1e-4 * np.abs(Ho).max())
        assert np.allclose(Hc, Hc.T)
    cuda_api.destroy(hc)
    oracle_api.destroy(ho)


def test_align_parity_c1(cuda_api, oracle_api, tum_calib, pair_c1):
    bgr_a, d_a, bgr_b, d_b, T_gt = pair_c1
    hc, ho = _both(cuda_api, oracle_api, tum_calib)
    for api, h in ((cuda_api, hc), (oracle_api, ho)):
        api.set_frame(h, 0, bgr_a, d_a)
        api.set_frame(h, 1, bgr_b, d_b)
    rc, _ = cuda_api.align(hc)
    ro, _ = oracle_api.align(ho)
    ang, dist = pose_error(rc.transform_np(), ro.transform_np())
    print("C1: iterations gpu/oracle", rc.iterations, ro.iterations, "pose diff", ang, dist)
    assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
    assert rc.iterations == ro.iterations and rc.A_nonzero == ro.A_nonzero
    ang, dist = pose_error(rc.transform_np(), T_gt)
    assert ang < 5e-3 and dist < 5e-3
    # prev_transform / accum_transform source (cvo.cpp:815-816): same bits as the oracle's
    assert np.array_equal(rc.last_iter_transform_np(), ro.last_iter_transform_np())
    assert not np.array_equal(rc.last_iter_transform_np(), rc.transform_np())
    # state persists: a second align of the same object starts from R, T, ell left behind
    assert cuda_api.get_ell(hc) == pytest.approx(oracle_api.get_ell(ho))
    Rc, Tc = cuda_api.get_RT(hc)
    assert np.allclose(Rc, rc.R_np()) and np.allclose(Tc, rc.T_np())
    rc2, _ = cuda_api.align(hc)
    ro2, _ = oracle_api.align(ho)
    ang, dist = pose_error(rc2.transform_np(), ro2.transform_np())
    assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
    assert rc2.iterations == ro2.iterations
    cuda_api.destroy(hc)
    oracle_api.destroy(ho)


def test_align_fast_mode_same_basin(cuda_api, oracle_api, tum_calib, pair_c1):
    """exp_mode = 1 (MUFU ex2): same schedule for the first 10 iterations within 1e-4; free running
    it ends in the same basin as the oracle (the tail is chaotic)."""
    bgr_a, d_a, bgr_b, d_b, T_gt = pair_c1
    for max_iter, tol in ((10, 1e-4), (2000, 2e-3)):
        p = cuda_api.default_params()
        p.exp_mode = 1
        p.max_iter = max_iter
        q = oracle_api.default_params()
        q.max_iter = max_iter
        hc, ho = cuda_api.create(tum_calib, p), oracle_api.create(tum_calib, q)
        for api, h in ((cuda_api, hc), (oracle_api, ho)):
            api.set_frame(h, 0, bgr_a, d_a)
            api.set_frame(h, 1, bgr_b, d_b)
        rc, _ = cuda_api.align(hc)
        ro, _ = oracle_api.align(ho)
        ang, dist = pose_error(rc.transform_np(), ro.transform_np())
        print("fast mode max_iter", max_iter, "iterations", rc.iterations, ro.iterations, "pose diff", ang, dist)
        assert ang < tol and dist < tol
        cuda_api.destroy(hc)
        oracle_api.destroy(ho)


def test_align_eth3d_large_ell(cuda_api, oracle_api):
    """C4: 739x458 pair, larger motion, default and ell_init = 0.25 (wide cutoff, many neighbours)."""
    from cvo_slam_b200 import capi, synth
    cal = capi.ETH3D_CALIB()
    a, da, b, db, T_gt = synth.make_pair(4, cal, w=739, h=458, rot_deg=2.0, trans=(0.04, -0.02, 0.03))
    for ell in (None, 0.25):
        p = cuda_api.default_params()
        if ell:
            p.ell_init = ell
        hc, ho = _both(cuda_api, oracle_api, cal, p)
        for api, h in ((cuda_api, hc), (oracle_api, ho)):
            api.set_frame(h, 0, a, da)
            api.set_frame(h, 1, b, db)
        rc, _ = cuda_api.align(hc)
        ro, _ = oracle_api.align(ho)
        assert rc.status == 0
        ang, dist = pose_error(rc.transform_np(), ro.transform_np())
        print("C4 ell", ell, "iterations", rc.iterations, ro.iterations, "pose diff", ang, dist)
        assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
        assert rc.iterations == ro.iterations
        cuda_api.destroy(hc)
        oracle_api.destroy(ho)



def test_align_c4_specified_large_motion(cuda_api, oracle_api):
    """C4 exactly as SURVEY 8(d) / BASELINE.md state it: 739x458, ETH3D intrinsics, 0.15 m / 8 deg offset,
    default parameters and ell_init = 0.25.  From this far the reference's schedule (ell drops to 0.03 after
    20 iterations) leaves the basin of the true pose: the oracle runs 700-900 iterations and stops ~5 deg /
    11 cm from the ground truth.  What is gateable, and gated: the CUDA path walks the SAME long trajectory
    (iteration count, final pose <= 1e-4, non-zeros, status 0); many neighbour-list rebuilds, the widest
    cutoff and the largest lists of the suite are on this path."""
    from cvo_slam_b200 import capi, synth
    cal = capi.ETH3D_CALIB()
    t = np.array([0.10, -0.05, 0.10])
    t = t / np.linalg.norm(t) * 0.15
    a, da, b, db, T_gt = synth.make_pair(4, cal, w=739, h=458, rot_deg=8.0, trans=tuple(t))
    for ell in (None, 0.25):
        p = cuda_api.default_params()
        if ell:
            p.ell_init = ell
        hc, ho = _both(cuda_api, oracle_api, cal, p)
        for api, h in ((cuda_api, hc), (oracle_api, ho)):
            api.set_frame(h, 0, a, da)
            api.set_frame(h, 1, b, db)
        rc, recs_c = cuda_api.align(hc, trace_cap=2000)
        ro, recs_o = oracle_api.align(ho, trace_cap=2000)
        ang, dist = pose_error(rc.transform_np(), ro.transform_np())
        ang_gt, dist_gt = pose_error(ro.transform_np(), T_gt)
        print("C4 8deg/15cm ell", ell, "iterations", rc.iterations, ro.iterations, "pose diff", ang, dist,
              "oracle vs ground truth", ang_gt, dist_gt, "max nnz", max(r["nnz"] for r in recs_o))
        assert rc.status == 0
        assert rc.iterations == ro.iterations and rc.A_nonzero == ro.A_nonzero
        assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
        n = min(len(recs_c), len(recs_o))
        assert np.array_equal([r["nnz"] for r in recs_c[:n]], [r["nnz"] for r in recs_o[:n]])
        # inner products at the final state, same transform on both sides
        T = ro.transform_np()
        for sa, Ta, sb in ((1, None, 0), (1, T, 0)):
            vc, nc = cuda_api.inner_product(hc, sa, Ta, sb)
            vo, no = oracle_api.inner_product(ho, sa, Ta, sb)
            assert nc == no and vc == pytest.approx(vo, rel=INNER_RTOL)
        cuda_api.destroy(hc)
        oracle_api.destroy(ho)


def test_edge_cases(cuda_api, oracle_api, tum_calib):
    from cvo_slam_b200.capi import CvoError
    h = cuda_api.create(tum_calib)
    with pytest.raises(CvoError):
        cuda_api.align(h)                      # "cvo not initialized !" -> CVO_ERR_NOT_INIT
    rng = np.random.default_rng(7)
    a = rng.uniform(0, 1, (50, 3)).astype(np.float32)
    f = rng.uniform(0, 255, (50, 5)).astype(np.float32)
    cuda_api.set_cloud(h, 0, a, f)
    cuda_api.set_cloud(h, 1, a + 10.0, f)      # far apart: empty neighbourhoods
    res, recs = cuda_api.align(h, trace_cap=4)
    assert res.iterations == 1 and res.iter == 0 and res.A_nonzero == 0
    assert recs[0]["step"] == pytest.approx(0.2)
    assert np.array_equal(res.transform_np(), np.eye(4, dtype=np.float32))
    v, n = cuda_api.inner_product(h, 1, None, 0)
    assert v == 0 and n == 1
    Hm, inl = cuda_api.hessian(h, 1, None, 0)
    assert inl == 0 and np.array_equal(Hm, np.eye(6))
    # empty cloud
    cuda_api.set_cloud(h, 1, np.zeros((0, 3), np.float32), np.zeros((0, 5), np.float32))
    res, _ = cuda_api.align(h)
    assert res.A_nonzero == 0 and res.iterations == 1
    # ragged sizes, identical clouds: zero flow at the optimum of <x, x>? not zero, but finite
    cuda_api.set_cloud(h, 1, a[:17], f[:17])
    res, _ = cuda_api.align(h)
    assert np.isfinite(res.transform_np()).all()
    # slot moves
    cuda_api.slot_move(h, 2, 1)
    assert cuda_api.slot_size(h, 1) == -1 and cuda_api.slot_size(h, 2) == 17
    cuda_api.destroy(h)


def test_batch_matches_single_and_oracle(cuda_api, oracle_api, tum_calib):
    import ctypes as C
    from cvo_slam_b200 import batch as B, synth
    scene = synth.make_scene(5)
    rng = np.random.default_rng(5)
    frames = []
    poses = [synth.pose()] + [synth.pose(rng.normal(0, 6e-3, 3), rng.normal(0, 8e-3, 3)) for _ in range(3)]
    for k, P in enumerate(poses):
        frames.append(synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=50 + k)))
    bgr = np.stack([f[0] for f in frames])
    dep = np.stack([f[1] for f in frames])
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (2, 3), (3, 1), (1, 0)]
    bt = B.Batch(tum_calib, max_frames=4, max_pairs=len(pairs), width=640, height=480)
    bt.set_frames(bgr, dep)
    sizes = [bt.frame_size(k) for k in range(4)]
    res = bt.align(pairs)
    vals, nums = bt.inner_product(pairs, res)
    assert all(r["status"] == 0 for r in res)
    for (fi, mi), r, v in zip(pairs, res, vals):
        ho = oracle_api.create(tum_calib)
        oracle_api.set_frame(ho, 0, *frames[fi])
        oracle_api.set_frame(ho, 1, *frames[mi])
        assert oracle_api.slot_size(ho, 0) == sizes[fi]
        ro, _ = oracle_api.align(ho)
        ang, dist = pose_error(r["transform"].reshape(4, 4), ro.transform_np())
        assert ang < POSE_TOL_RAD and dist < POSE_TOL_M, (fi, mi, ang, dist)
        gt = synth.relative_transform(poses[fi], poses[mi])
        ang, dist = pose_error(r["transform"].reshape(4, 4), gt)
        assert ang < 6e-3 and dist < 6e-3
        vo, _ = oracle_api.inner_product(ho, 1, ro.transform_np(), 0)
        assert v == pytest.approx(vo, rel=INNER_RTOL)
        assert r["iterations"] == ro.iterations
        oracle_api.destroy(ho)
    # the batch path and the handle path run the same kernel: identical bits
    hc = cuda_api.create(tum_calib)
    cuda_api.set_frame(hc, 0, *frames[0])
    cuda_api.set_frame(hc, 1, *frames[1])
    rs, _ = cuda_api.align(hc)
    assert np.array_equal(rs.transform_np().reshape(-1), res[0]["transform"])
    cuda_api.destroy(hc)
    st = bt.stats()
    assert st["launches"] > 0 and st["evals"] > 0
    bt.close()


def test_batch_pipelined_uploads_match_plain(cuda_api, tum_calib):
    """cvo_batch_set_frames only enqueues: the frames of the next step can be uploaded into a second range of
    the arena before the current step is aligned.  Every step must give the bits of a plain
    upload-then-align batch, also when a range is overwritten right after it was aligned out of."""
    from cvo_slam_b200 import batch as B, synth
    scene = synth.make_scene(8)
    rng = np.random.default_rng(8)
    n = 3
    sets = []
    for s_ in range(3):   # three different sets of n frames
        poses = [synth.pose()] + [synth.pose(rng.normal(0, 6e-3, 3), rng.normal(0, 8e-3, 3)) for _ in range(n - 1)]
        fr = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=100 + 10 * s_ + k)) for k, P in enumerate(poses)]
        sets.append((np.ascontiguousarray(np.stack([f[0] for f in fr])), np.ascontiguousarray(np.stack([f[1] for f in fr]))))
    pairs = [(0, 1), (0, 2), (1, 2)]
    plain = []
    bt = B.Batch(tum_calib, max_frames=n, max_pairs=len(pairs), width=640, height=480)
    for bgr, dep in sets:
        bt.set_frames(bgr, dep)
        r = bt.align(pairs)
        v, c = bt.inner_product(pairs, r)
        plain.append((r["transform"].copy(), r["iterations"].copy(), np.array(v), np.array(c)))
    bt.close()
    assert not np.array_equal(plain[0][0], plain[1][0])
    bt = B.Batch(tum_calib, max_frames=2 * n, max_pairs=len(pairs), width=640, height=480)
    bank = [pairs, [(f + n, m + n) for f, m in pairs]]
    bt.set_frames(*sets[0], first=0)
    for k in range(3):
        if k + 1 < 3:
            bt.set_frames(*sets[k + 1], first=((k + 1) & 1) * n)   # set 2 overwrites the range set 0 was just aligned out of
        r = bt.align(bank[k & 1])
        v, c = bt.inner_product(bank[k & 1], r)
        assert all(x == 0 for x in r["status"])
        assert np.array_equal(r["transform"], plain[k][0]) and np.array_equal(r["iterations"], plain[k][1]), k
        assert np.array_equal(np.array(v), plain[k][2]) and np.array_equal(np.array(c), plain[k][3]), k
        assert all(bt.frame_size((k & 1) * n + j) > 1000 for j in range(n))
    bt.close()


def test_list_filter_path_is_exercised(cuda_api, tum_calib, pair_c1):
    """The C1 alignment changes its length scale three times; with the default skin at least one of the new
    neighbour lists is derived by filtering the current one — the parity tests above therefore
    cover that path.  Counters: cvo_handle_phase_cycles[6] = filters, [7] = grid searches."""
    a, da, b, db, _ = pair_c1
    for mode in (0, 1):
        p = cuda_api.default_params()
        p.exp_mode = mode
        h = cuda_api.create(tum_calib, p)
        cuda_api.set_frame(h, 0, a, da)
        cuda_api.set_frame(h, 1, b, db)
        res, _ = cuda_api.align(h)
        ph = cuda_api.phase_cycles(h)
        csize = 16   # a handle runs a cluster: every CTA counts its own list constructions
        assert res.status == 0
        assert ph["filters"] >= 1 and ph["rebuilds"] >= 1, ph
        assert (ph["filters"] + ph["rebuilds"]) <= 12 * csize, ph
        cuda_api.destroy(h)


def test_tracking_sequence_parity(cuda_api, oracle_api, tum_calib):
    """C2 in miniature: the LocalTracker call pattern over 6 frames, two cvo objects with
    persistent R/T/ell, GPU vs oracle."""
    from cvo_slam_b200 import cvo as cvo_mod, synth
    scene = synth.make_scene(2)
    poses = synth.trajectory(6, 2)
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=20 + k))
              for k, P in enumerate(poses)]
    out_c = cvo_mod.track_sequence(frames, tum_calib, api=cuda_api)
    out_o = cvo_mod.track_sequence(frames, tum_calib, api=oracle_api)
    for k, (c, o) in enumerate(zip(out_c, out_o)):
        for key in ("odometry", "keyframe"):
            ang, dist = pose_error(c[key], o[key])
            assert ang < POSE_TOL_RAD and dist < POSE_TOL_M, (k, key, ang, dist)
        for key in ("inn_pre", "inn_fixed_pcd", "inn_moving_pcd"):
            assert c["r_odometry"][key].value == pytest.approx(o["r_odometry"][key].value, rel=INNER_RTOL)
        assert c["r_odometry"]["inn_post"].value == pytest.approx(o["r_odometry"]["inn_post"].value, rel=INNER_RTOL)
        assert c["r_odometry"]["cos_angle"] == pytest.approx(o["r_odometry"]["cos_angle"], rel=INNER_RTOL)
        Hc, Ho = c["r_odometry"]["post_hessian"], o["r_odometry"]["post_hessian"]
        assert np.allclose(Hc, Ho, rtol=0, atol=1e-4 * np.abs(Ho).max())



def test_tracking_sequence_c2_300_frames(cuda_api, oracle_api, tum_calib):
    """C2 at full length (BASELINE configs[1]): the 300-frame synthetic sequence through the LocalTracker
    call pattern, CUDA path vs oracle.  Every odometry and keyframe pose within 1e-4 rad / 1e-4 m of the
    oracle's, same iteration counts, and every inner product of compute_innerproduct within 1e-4 relative.
    State (R, T, ell, the three cloud slots) persists over the whole chain, so one diverging frame would
    show in all later ones."""
    from cvo_slam_b200 import cvo as cvo_mod, synth
    n_frames = 300
    scene = synth.make_scene(2)
    poses = synth.trajectory(n_frames, 2)
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=20 + k, device="cuda:0"))
              for k, P in enumerate(poses)]
    out_c = cvo_mod.track_sequence(frames, tum_calib, api=cuda_api)
    out_o = cvo_mod.track_sequence(frames, tum_calib, api=oracle_api)
    assert len(out_c) == len(out_o) == n_frames - 1
    worst = [0.0, 0.0, 0.0]
    for k, (c, o) in enumerate(zip(out_c, out_o)):
        for key in ("odometry", "keyframe"):
            ang, dist = pose_error(c[key], o[key])
            worst[0], worst[1] = max(worst[0], ang), max(worst[1], dist)
            assert ang < POSE_TOL_RAD and dist < POSE_TOL_M, (k, key, ang, dist)
        for rk in ("r_odometry", "r_keyframe"):
            for key in ("inn_pre", "inn_post", "inn_fixed_pcd", "inn_moving_pcd"):
                vc, vo = float(c[rk][key].value), float(o[rk][key].value)
                worst[2] = max(worst[2], abs(vc - vo) / max(abs(vo), 1e-30))
                assert vc == pytest.approx(vo, rel=INNER_RTOL), (k, rk, key)
                assert c[rk][key].num == o[rk][key].num, (k, rk, key)
            assert float(c[rk]["cos_angle"]) == pytest.approx(float(o[rk]["cos_angle"]), rel=INNER_RTOL)
            assert c[rk]["inliers"] == o[rk]["inliers"]
            Hc, Ho = c[rk]["post_hessian"], o[rk]["post_hessian"]
            assert np.allclose(Hc, Ho, rtol=0, atol=1e-4 * np.abs(Ho).max()), (k, rk)
    gt_err = [pose_error(o["keyframe"], synth.relative_transform(poses[0], poses[k + 1])) for k, o in enumerate(out_c)]
    print("C2 300 frames: worst pose diff vs oracle", worst[0], worst[1], "worst inner-product rel diff", worst[2],
          "max keyframe error vs ground truth", max(e[0] for e in gt_err), max(e[1] for e in gt_err))


def test_dedup_front_end_matches_oracle(cuda_api, oracle_api, tum_calib):
    """the keyframe object adopts the odometry object's selection of the same image
    (cvo::set_pcd_from / match_keyframe_from over cvo_copy_cloud) — checked against the ORACLE driven the reference's
    way (two selections per frame), not against a second CUDA handle."""
    from cvo_slam_b200 import cvo as cvo_mod, synth
    scene = synth.make_scene(2)
    poses = synth.trajectory(10, 2)
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=20 + k))
              for k, P in enumerate(poses)]
    out_c = cvo_mod.track_sequence(frames, tum_calib, api=cuda_api, dedup=True)
    out_o = cvo_mod.track_sequence(frames, tum_calib, api=oracle_api)
    for k, (c, o) in enumerate(zip(out_c, out_o)):
        for key in ("odometry", "keyframe"):
            ang, dist = pose_error(c[key], o[key])
            assert ang < POSE_TOL_RAD and dist < POSE_TOL_M, (k, key, ang, dist)
        for rk in ("r_odometry", "r_keyframe"):
            for key in ("inn_pre", "inn_post", "inn_fixed_pcd", "inn_moving_pcd"):
                assert c[rk][key].num == o[rk][key].num
                assert float(c[rk][key].value) == pytest.approx(float(o[rk][key].value), rel=INNER_RTOL)
    # and bit-identical to the CUDA path that selects twice
    out_2 = cvo_mod.track_sequence(frames, tum_calib, api=cuda_api)
    assert all(np.array_equal(a["keyframe"], b["keyframe"]) for a, b in zip(out_c, out_2))


def test_cpp_dropin_matches_python_path(cuda_api, tum_calib, pair_c1, tmp_path):
    """The C++ drop-in class (include/cvo.hpp) driven like LocalTracker gives the same bits as the
    ctypes path: same library, same kernels."""
    import os
    import subprocess
    from test_host_logic import _build_dropin
    from cvo_slam_b200 import cvo as cvo_mod
    bgr_a, d_a, bgr_b, d_b, _ = pair_c1
    exe = _build_dropin(tmp_path)
    files = []
    for name, arr in (("a_bgr", bgr_a), ("a_d", d_a), ("b_bgr", bgr_b), ("b_d", d_b)):
        p = os.path.join(str(tmp_path), name + ".raw")
        arr.tofile(p)
        files.append(p)
    calib = os.path.join(str(tmp_path), "calib.yaml")
    with open(calib, "w") as f:
        f.write("%YAML:1.0\nCamera.fx: 517.306408\nCamera.fy: 516.469215\nCamera.cx: 318.643040\n"
                "Camera.cy: 255.313989\nDepthMapFactor: 5000.0\n")
    out = subprocess.run([exe, calib] + files + ["640", "480"], capture_output=True, text=True, check=True).stdout
    assert "cvo not initialized !" in out
    lines = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l and l.split()[0] in ("N", "T", "inn")}
    c = cvo_mod.Cvo(tum_calib, api=cuda_api)
    c.set_pcd(bgr_a, d_a)
    T = c.match_odometry(bgr_b, d_b)
    r = c.compute_innerproduct(T.astype(np.float32))
    assert [int(lines["N"][0]), int(lines["N"][1])] == list(c.get_fixed_and_moving_number())
    assert int(lines["N"][3]) == c.get_A_nonzero() and int(lines["N"][5]) == c.get_iteration_number()
    Tc = np.array([float(x) for x in lines["T"]], np.float32).reshape(4, 4)
    assert np.array_equal(Tc, c.transform)
    inn = [float(x) for x in lines["inn"][:4]]
    ref = [r["inn_pre"].value, r["inn_post"].value, r["inn_fixed_pcd"].value, r["inn_moving_pcd"].value]
    assert np.allclose(inn, ref, rtol=1e-6)
    fip = [l.split() for l in out.splitlines() if l.startswith("fip ")][0]
    assert float(fip[1]) == pytest.approx(float(fip[4]), rel=1e-6) and fip[2] == fip[5]   # wrapper == slot path
    assert int(fip[7]) == int(fip[2])                                                  # same pair set
    # the reference's getter returns what set_pcd cached, also after the clouds have moved (cvo.cpp:370-371, 578-582)
    c.update_fixed_pcd()
    assert "after update_fixed_pcd N %d %d" % c.get_fixed_and_moving_number() in out
    c.close()


def test_dense_c3_parity(cuda_api, oracle_plain, tum_calib):
    """C3: dense selection (num_want = 60000 -> pot 1, ~18 k points per frame, ~1.4 M non-zeros at
    ell = 0.15).  Exercises the large-cloud scratch sizing and the cluster path."""
    from cvo_slam_b200 import synth
    a, da, b, db, T_gt = synth.make_pair(3, tum_calib, high_gradient=True, rot_deg=0.8, trans=(0.015, -0.01, 0.012))
    p = cuda_api.default_params()
    p.num_want = 60000
    hc, ho = cuda_api.create(tum_calib, p), oracle_plain.create(tum_calib, p, search=0)
    for api, h in ((cuda_api, hc), (oracle_plain, ho)):
        api.set_frame(h, 0, a, da)
        api.set_frame(h, 1, b, db)
    n = cuda_api.slot_size(hc, 0)
    assert n == oracle_plain.slot_size(ho, 0) and n > 15000
    I, z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    rc = cuda_api.iteration_at(hc, I, z, 0.15)
    ro = oracle_plain.iteration_at(ho, I, z, 0.15)
    assert rc["nnz"] == ro["nnz"] > 1000000
    assert np.array_equal(rc["omega"], ro["omega"]) and np.array_equal(rc["v"], ro["v"])
    assert rc["step"] == pytest.approx(ro["step"], rel=1e-6)
    res_c, _ = cuda_api.align(hc)
    res_o, _ = oracle_plain.align(ho)
    assert res_c.status == 0
    ang, dist = pose_error(res_c.transform_np(), res_o.transform_np())
    print("C3 dense: N", n, "iterations", res_c.iterations, res_o.iterations, "pose diff", ang, dist)
    assert res_c.iterations == res_o.iterations
    assert ang < POSE_TOL_RAD and dist < POSE_TOL_M
    cuda_api.destroy(hc)
    oracle_plain.destroy(ho)


def test_copy_cloud_between_handles(cuda_api, tum_calib, pair_c1):
    """cvo_copy_cloud: select once, share the device cloud with a second object."""
    bgr_a, d_a, bgr_b, d_b, _ = pair_c1
    h1, h2 = cuda_api.create(tum_calib), cuda_api.create(tum_calib)
    cuda_api.set_frame(h1, 0, bgr_a, d_a)
    cuda_api.set_frame(h1, 1, bgr_b, d_b)
    cuda_api.copy_cloud(h2, 0, h1, 0)
    cuda_api.copy_cloud(h2, 1, h1, 1)
    for slot in (0, 1):
        p1, f1 = cuda_api.get_cloud(h1, slot)
        p2, f2 = cuda_api.get_cloud(h2, slot)
        assert np.array_equal(p1, p2) and np.array_equal(f1, f2)
        assert np.array_equal(cuda_api.get_selected_points(h1, slot), cuda_api.get_selected_points(h2, slot))
    r1, _ = cuda_api.align(h1)
    r2, _ = cuda_api.align(h2)
    assert np.array_equal(r1.transform_np(), r2.transform_np()) and r1.iterations == r2.iterations
    cuda_api.destroy(h1)
    cuda_api.destroy(h2)


def test_two_handles_from_two_threads(cuda_api, tum_calib, pair_c1):
    """Distinct handles are independent (one stream each) and may be driven concurrently, like the
    loop-closure object on the optimisation thread (keyframe_graph.cpp:151-154)."""
    import threading
    bgr_a, d_a, bgr_b, d_b, _ = pair_c1
    out = {}

    def work(name):
        h = cuda_api.create(tum_calib)
        cuda_api.set_frame(h, 0, bgr_a, d_a)
        cuda_api.set_frame(h, 1, bgr_b, d_b)
        for _ in range(3):
            cuda_api.set_RT(h, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
            cuda_api.set_ell(h, 0.15)
            r, _ = cuda_api.align(h)
        out[name] = (r.transform_np(), r.iterations, cuda_api.inner_product(h, 1, r.transform_np(), 0))
        cuda_api.destroy(h)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(out) == 3
    for k in (1, 2):
        assert np.array_equal(out[k][0], out[0][0]) and out[k][1] == out[0][1]
        assert out[k][2][1] == out[0][2][1]


def test_tiny_and_degenerate_clouds(cuda_api, oracle_plain, tum_calib):
    """one-point clouds, coincident points, identical clouds: same results as the oracle."""
    rng = np.random.default_rng(9)
    f = rng.integers(0, 256, (64, 5)).astype(np.float32)
    base = rng.uniform(-0.3, 0.3, (64, 3)).astype(np.float32)
    base[:, 2] += 1.5
    cases = [(base[:1], f[:1], base[:1] + 0.01, f[:1]),                 # single points
             (base, f, base.copy(), f),                                  # identical clouds
             (np.repeat(base[:4], 16, 0), f, base + 0.005, f),           # coincident fixed points
             (base[:33], f[:33], base[:7] + 0.004, f[:7])]               # ragged sizes
    for pa, fa_, pb, fb in cases:
        hc, ho = cuda_api.create(tum_calib), oracle_plain.create(tum_calib, search=1)
        for api, h in ((cuda_api, hc), (oracle_plain, ho)):
            api.set_cloud(h, 0, pa, fa_)
            api.set_cloud(h, 1, pb, fb)
        rc, _ = cuda_api.align(hc)
        ro, _ = oracle_plain.align(ho)
        assert rc.status == 0
        assert rc.iterations == ro.iterations and rc.A_nonzero == ro.A_nonzero
        assert np.allclose(rc.transform_np(), ro.transform_np(), atol=1e-6)
        vc, nc = cuda_api.inner_product(hc, 1, None, 0)
        vo, no = oracle_plain.inner_product(ho, 1, None, 0)
        assert nc == no and vc == pytest.approx(vo, rel=1e-4, abs=1e-12)
        cuda_api.destroy(hc)
        oracle_plain.destroy(ho)


def test_invalid_arguments_return_codes(cuda_api, tum_calib):
    import ctypes as C
    from cvo_slam_b200.capi import CvoError
    lib = cuda_api.lib
    h = cuda_api.create(tum_calib)
    assert lib.cvo_slot_move(h, 0, 7) == -1
    assert lib.cvo_set_frame(h, 0, None, 0, None, 0, 640, 480) == -1
    n = C.c_int(0)
    assert lib.cvo_slot_size(h, 1, C.byref(n)) == -3          # empty slot
    with pytest.raises(CvoError):
        cuda_api.inner_product(h, 1, None, 0)
    too_small = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(CvoError):
        cuda_api.set_frame(h, 0, too_small, np.zeros((32, 32), np.uint16))
    # loop-closure verification: clouds missing -> not initialised; null pointers -> invalid
    I = np.eye(4, dtype=np.float32)
    with pytest.raises(CvoError):
        cuda_api.compute_innerproduct_lc(h, I, I, I, I)
    assert lib.cvo_compute_innerproduct_lc(h, None, None, None, None, None) == -1
    assert lib.cvo_batch_verify_lc(None, 1, None, None, None, None, None, None) == -1
    cuda_api.destroy(h)


def _perturbed(T, rot_deg, trans, seed):
    """T composed with a small rigid perturbation (stand-ins for the PnP-RANSAC / SVD priors)."""
    from cvo_slam_b200 import synth
    rng = np.random.default_rng(seed)
    P = synth.pose(rng.normal(0, np.deg2rad(rot_deg), 3), rng.normal(0, trans, 3)).astype(np.float32)
    return (np.asarray(T, np.float32) @ P).astype(np.float32)


@pytest.mark.gpu
def test_compute_innerproduct_lc_fused_matches_oracle(cuda_api, oracle_api, tum_calib, pair_c1):
    """cvo::compute_innerproduct_lc (cvo.cpp:505-561): the one-launch CUDA entry against the oracle's
    eight separate queries, same transforms and ell on both sides; and the accept rule of
    keyframe_graph.cpp:711-712."""
    from cvo_slam_b200 import cvo as cvo_mod
    bgr_a, d_a, bgr_b, d_b, T_gt = pair_c1
    out = {}
    for name, api in (("cuda", cuda_api), ("oracle", oracle_api)):
        c = cvo_mod.Cvo(tum_calib, api=api)
        c.set_pcd(bgr_a, d_a)
        T = c.match_keyframe(bgr_b, d_b).astype(np.float32)
        prior, lc_prior, lc_prior2 = _perturbed(T, 0.6, 8e-3, 1), _perturbed(T, 0.3, 4e-3, 2), _perturbed(T, 0.2, 3e-3, 3)
        out[name] = (T, c.compute_innerproduct_lc(prior, lc_prior, lc_prior2, T))
        c.close()
    (Tc, rc), (To, ro) = out["cuda"], out["oracle"]
    assert np.array_equal(Tc, To)   # bit-faithful alignment: identical transforms feed both sides
    for k in ("inn_prior", "inn_lc_prior", "inn_lc_pre", "inn_lc_post", "inn_fixed_pcd", "inn_moving_pcd"):
        assert rc[k].num == ro[k].num, k
        assert rc[k].value == pytest.approx(ro[k].value, rel=INNER_RTOL), k
    assert rc["inliers_svd"] == ro["inliers_svd"] and rc["inliers_pnpransac"] == ro["inliers_pnpransac"]
    assert float(rc["cos_angle"]) == pytest.approx(float(ro["cos_angle"]), rel=1e-4)
    Hc, Ho = rc["post_hessian"], ro["post_hessian"]
    assert np.allclose(Hc, Ho, rtol=0, atol=1e-4 * np.abs(Ho).max())
    # the aligned transform has the largest inner product: the candidate is accepted
    assert rc["accept"] is True
    assert rc["inn_lc_post"].value > max(rc["inn_lc_pre"].value, rc["inn_lc_prior"].value, rc["inn_prior"].value)


@pytest.mark.gpu
def test_batch_verify_lc_matches_handle_path(cuda_api, tum_calib):
    """cvo_batch_verify_lc == per-pair cvo_compute_innerproduct_lc on a handle (same kernel, same
    queries), including shared self inner products and a rejected candidate."""
    from cvo_slam_b200 import batch as B, synth
    scene = synth.make_scene(7)
    rng = np.random.default_rng(7)
    poses = [synth.pose()] + [synth.pose(rng.normal(0, 6e-3, 3), rng.normal(0, 8e-3, 3)) for _ in range(2)]
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=70 + k)) for k, P in enumerate(poses)]
    pairs = [(0, 1), (0, 2), (1, 2), (2, 0)]
    bt = B.Batch(tum_calib, max_frames=3, max_pairs=len(pairs), width=640, height=480)
    bt.set_frames(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]))
    res = bt.align(pairs)
    T = res["transform"].reshape(-1, 4, 4)
    prior = np.stack([_perturbed(T[i], 0.6, 8e-3, 10 + i) for i in range(len(pairs))])
    lc_prior = np.stack([_perturbed(T[i], 0.3, 4e-3, 20 + i) for i in range(len(pairs))])
    lc_prior2 = np.stack([_perturbed(T[i], 0.2, 3e-3, 30 + i) for i in range(len(pairs))])
    lc_prior[3] = T[3]   # a prior as good as the CVO result: inn_post <= inn_lc_prior -> reject
    out = bt.verify_lc(pairs, res, prior, lc_prior, lc_prior2)
    for i, (fi, mi) in enumerate(pairs):
        h = cuda_api.create(tum_calib)
        cuda_api.set_frame(h, 0, *frames[fi])
        cuda_api.set_frame(h, 1, *frames[mi])
        cuda_api.set_ell(h, float(res[i]["ell"]))
        r = cuda_api.compute_innerproduct_lc(h, prior[i], lc_prior[i], lc_prior2[i], T[i])
        assert np.array_equal(np.array(r.value[:], np.float32), out[i]["value"]), i
        assert list(r.num[:]) == list(out[i]["num"])
        assert np.array_equal(np.array(r.post_hessian[:]), out[i]["post_hessian"])
        assert (r.inliers_svd, r.inliers_pnpransac, r.accept) == (out[i]["inliers_svd"], out[i]["inliers_pnpransac"], out[i]["accept"])
        assert r.cos_angle == out[i]["cos_angle"]
        cuda_api.destroy(h)
    assert list(out["accept"]) == [1, 1, 1, 0]
    bt.close()


def test_batch_shape_properties_512_pairs(cuda_api, tum_calib):
    """The bench's workload shape at 1/16 of its size (64 keyframes x 8 partners = 512 pairs, priors via
    reset_initial), checked through properties that do not need the oracle: every pair converges to the
    rendered ground truth, a second run is bit-identical (no result depends on the order in which
    CTAs pull pairs or warps append list entries), the (moving, fixed) swap gives the inverse pose
    within the convergence basin, and the handle path reproduces sampled pairs to the bit."""
    import importlib.util
    import os
    from conftest import ROOT
    from cvo_slam_b200 import batch as B
    spec = importlib.util.spec_from_file_location("cvo_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n_frames, partners, seed = 64, 8, 1000
    pairs = bench.pair_list(n_frames, partners)
    poses = bench.keyframe_poses(n_frames, seed)
    R0, T0, gts = bench.pair_priors(pairs, poses, seed)
    bgr_d, dep_d = bench.render_frames(list(range(n_frames)), poses, seed, "cuda:0")
    prm = cuda_api.default_params()
    bt = B.Batch(tum_calib, prm, max_frames=n_frames, max_pairs=len(pairs), width=640, height=480)
    bt.set_frames_ptr(bgr_d.data_ptr(), dep_d.data_ptr(), n_frames, device=True)
    desc = bt.make_pairs(pairs, R0.reshape(-1, 3, 3), T0, prm.ell_init)
    res = bt.align(desc)
    assert (res["status"] == 0).all()
    errs = np.array([pose_error(r["transform"].reshape(4, 4), gt) for r, gt in zip(res, gts)])
    assert np.median(errs[:, 1]) < 2e-3 and errs[:, 0].max() < 2e-2 and errs[:, 1].max() < 2e-2, errs.max(axis=0)
    res2 = bt.align(desc)
    assert np.array_equal(res["transform"], res2["transform"]) and np.array_equal(res["iterations"], res2["iterations"])
    # swapped roles, fresh prior = inverse of the original prior: the result is the inverse pose up to
    # the basin size (the two problems are different discretisations of the same registration)
    sw = [(m, f) for f, m in pairs[:64]]
    Rt = np.tile(np.eye(4), (64, 1, 1))
    Rt[:, :3, :3] = R0.reshape(-1, 3, 3)[:64]
    Rt[:, :3, 3] = T0[:64]
    Rti = np.linalg.inv(Rt)
    res_sw = bt.align(bt.make_pairs(sw, Rti[:, :3, :3].astype(np.float32), Rti[:, :3, 3].astype(np.float32), prm.ell_init))
    for r, rs in zip(res[:64], res_sw):
        ang, dist = pose_error(np.linalg.inv(rs["transform"].reshape(4, 4).astype(np.float64)), r["transform"].reshape(4, 4))
        assert ang < 1e-2 and dist < 1e-2, (ang, dist)
    # handle path == batch path for sampled pairs (same kernel, cluster of CTAs instead of one CTA)
    frames = [(bgr_d[k].cpu().numpy(), dep_d[k].cpu().numpy().view(np.uint16)) for k in range(n_frames)]
    for k in (0, 137, 511):
        f, m = pairs[k]
        h = cuda_api.create(tum_calib, prm)
        cuda_api.set_frame(h, 0, *frames[f])
        cuda_api.set_frame(h, 1, *frames[m])
        cuda_api.set_RT(h, R0.reshape(-1, 3, 3)[k], T0[k])
        cuda_api.set_ell(h, prm.ell_init)
        rs, _ = cuda_api.align(h)
        assert np.array_equal(rs.transform_np().reshape(-1), res[k]["transform"]), k
        assert rs.iterations == res[k]["iterations"]
        cuda_api.destroy(h)
    bt.close()


def test_cpp_dropin_sequence_matches_python_mirror(cuda_api, tum_calib, tmp_path):
    """LocalTracker's per-frame pattern (two cvo objects, reset_initial, keyframe tracking) through the
    drop-in C++ class (scripts/seq_dropin.cpp over include/cvo.hpp) ends on the same bits as the Python
    mirror of the class: the host logic of cvo.cpp:578-618 is the same arithmetic in both."""
    import importlib.util
    import os
    from conftest import ROOT
    from cvo_slam_b200 import cvo as cvo_mod, synth
    spec = importlib.util.spec_from_file_location("cvo_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    scene = synth.make_scene(2)
    poses = synth.trajectory(8, 2)
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=20 + k)) for k, P in enumerate(poses)]
    out = cvo_mod.track_sequence(frames, tum_calib, api=cuda_api)
    cpp = bench.cpp_sequence(frames, out[-1]["keyframe"])
    assert cpp is not None, "g++ or the runner failed"
    assert cpp["same_bits"]


def test_selected_points_device_pointer(cuda_api, tum_calib, pair_c1):
    """cvo_get_selected_points_device hands out the pixels the host getter copies (ORB-side consumer)."""
    import torch
    bgr_a, d_a, _, _, _ = pair_c1
    h = cuda_api.create(tum_calib)
    cuda_api.set_frame(h, 0, bgr_a, d_a)
    host = cuda_api.get_selected_points(h, 0)
    ptr, n = cuda_api.get_selected_points_device(h, 0)
    assert n == len(host) and ptr
    # read the device memory back through the CUDA runtime torch ships (same primary context)
    import ctypes as C
    import glob
    import os
    libs = sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")))
    assert libs, "libcudart not found next to torch"
    rt = C.CDLL(libs[0])
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    out = np.zeros((n, 2), np.float32)
    assert rt.cudaMemcpy(out.ctypes.data, ptr, n * 8, 2) == 0   # cudaMemcpyDeviceToHost
    assert np.array_equal(out, host)
    # the device gray image of the same frame (what Keyframe recomputes with cv::cvtColor, include/keyframe.h:34-55)
    gp, w, hh = C.c_void_p(), C.c_int(0), C.c_int(0)
    cuda_api.lib.cvo_get_gray_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    assert cuda_api.lib.cvo_get_gray_device(h, 0, C.byref(gp), C.byref(w), C.byref(hh)) == 0
    assert (w.value, hh.value) == (640, 480)
    gray = np.zeros((480, 640), np.uint8)
    assert rt.cudaMemcpy(gray.ctypes.data, gp, gray.size, 2) == 0
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(gray, cv2.cvtColor(bgr_a, cv2.COLOR_RGB2GRAY))
    assert cuda_api.lib.cvo_get_gray_device(h, 1, C.byref(gp), None, None) == -3   # nothing set on that slot last
    cuda_api.destroy(h)


def test_multi_device_entry_matches_single_batch(cuda_api, tum_calib):
    """cvo_multi_align (one list of host frames and pairs over n devices, contiguous blocks of pairs, one host
    thread per device) gives the bits of one cvo_batch on one GPU.  With a single GPU in the box the two
    "devices" are the same ordinal twice: the block split, the frame remapping and the threads are the same."""
    import torch
    from cvo_slam_b200 import batch as B, synth
    scene = synth.make_scene(8)
    rng = np.random.default_rng(8)
    poses = [synth.pose()] + [synth.pose(rng.normal(0, 6e-3, 3), rng.normal(0, 8e-3, 3)) for _ in range(5)]
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=80 + k)) for k, P in enumerate(poses)]
    bgr = np.stack([f[0] for f in frames])
    dep = np.stack([f[1] for f in frames])
    pairs = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 5), (5, 3), (5, 0), (2, 0)]
    bt = B.Batch(tum_calib, max_frames=6, max_pairs=len(pairs), width=640, height=480)
    bt.set_frames(bgr, dep)
    desc = bt.make_pairs(pairs)
    ref = bt.align(desc)
    rv, rn = bt.inner_product(desc, ref)
    bt.close()
    have = torch.cuda.device_count()
    for devices in ([0], [0, 1 % have], [0, 1 % have, 2 % have]):
        mb = B.MultiBatch(tum_calib, n_devices=len(devices), devices=devices, max_frames=6, max_pairs=len(pairs))
        res, vals, nums = mb.align(bgr, dep, desc)
        sh = mb.last_shares()
        assert sum(sh["pairs"]) == len(pairs) and max(sh["frames"]) <= 6
        assert np.array_equal(res["transform"], ref["transform"]) and np.array_equal(res["iterations"], ref["iterations"])
        assert np.array_equal(res["last_iter_transform"], ref["last_iter_transform"])
        assert np.array_equal(vals, rv) and np.array_equal(nums, rn)
        mb.close()
    # invalid arguments
    mb = B.MultiBatch(tum_calib, n_devices=1, max_frames=6, max_pairs=4)
    bad = desc[:2].copy()
    bad["moving_frame"][0] = 17
    assert mb.lib.cvo_multi_align(mb.m, 6, bgr.ctypes.data, dep.ctypes.data, 2, bad.ctypes.data,
                                  np.zeros(2, dtype=ref.dtype).ctypes.data, None, None) == -1
    mb.close()


def test_sharded_bench_two_gpus(tmp_path):
    """The strong-scaling arm of bench.py (one list of pairs sharded over the ranks, final gather on rank 0) on two
    GPUs: the gathered results equal the single-GPU results to the bit (bench.py asserts it inside the run)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--frames", "64", "--steps", "1",
           "--warmup", "1", "--no-cpu-baseline"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["scaling"] == "strong" and line["n_gpus"] == 2
    assert line["shard_check"]["bit_identical"] is True
    assert line["check"]["pairs_with_error_status"] == 0
    assert line["weak"]["scaling"] == "weak"
