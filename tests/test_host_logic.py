"""CPU-only checks of the product's host side: the C-ABI library loads and exports every
symbol include/cvo_b200.h declares (no compute without a GPU), struct layouts match, and
the host-side mirror of cvo::cvo drives the oracle correctly through the state shuffles
of cvo.cpp:578-618."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pose_error


def _header():
    return open(os.path.join(ROOT, "include", "cvo_b200.h")).read()


@pytest.fixture(scope="module")
def cuda_lib_path():
    from cvo_slam_b200 import build, capi
    build.build()
    return capi.LIB_PATH


def test_library_exports_every_declared_symbol(cuda_lib_path):
    lib = C.CDLL(cuda_lib_path)
    names = sorted(set(re.findall(r"\b(cvo_[a-z_A-Z0-9]+)\s*\(", _header())))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_library_is_sm100a_with_lineinfo(cuda_lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header(cuda_lib_path):
    """sizeof() as the C compiler sees the header vs the ctypes mirrors."""
    from cvo_slam_b200 import capi
    src = '#include "cvo_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",' \
          'sizeof(cvo_calib),sizeof(cvo_params),sizeof(cvo_align_result),sizeof(cvo_iter_record),sizeof(cvo_pair_desc));}'
    exe = "/tmp/_cvo_sizes"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(capi.Calib), C.sizeof(capi.Params), C.sizeof(capi.AlignResult),
            C.sizeof(capi.IterRecord), C.sizeof(capi.PairDesc)]
    assert got == want


def test_default_params_match_reference_constants(cuda_lib_path, oracle_plain):
    from cvo_slam_b200 import capi
    lib = C.CDLL(cuda_lib_path)
    p = capi.Params()
    lib.cvo_default_params(C.byref(p))
    q = oracle_plain.default_params()
    for name, _ in capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert (p.ell_init, p.max_iter, p.num_want, p.feature_type) == (np.float32(0.15), 2000, 3000, 1)


def test_random_pattern_restatement_matches_libc(cuda_lib_path, oracle_plain):
    """cvo_random_pattern restates glibc's TYPE_3 rand(); the oracle calls libc itself."""
    lib = C.CDLL(cuda_lib_path)
    n = 739 * 458
    out = np.zeros(n, np.uint8)
    assert lib.cvo_random_pattern(C.c_void_p(out.ctypes.data), n) == 0
    assert np.array_equal(out, oracle_plain.random_pattern(n))


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "cvo_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                for bad in ("import oracle", "from oracle", "cvo_oracle", "libcvo_oracle", "oracle/"):
                    assert bad not in txt, (f, bad)


def test_cvo_mirror_state_shuffles(oracle_plain, tum_calib):
    """update_fixed_pcd / update_previous_pcd / reset_keyframe / reset_initial semantics
    (cvo.cpp:578-618) through the Python mirror, driven by the oracle back end."""
    from cvo_slam_b200 import cvo as cvo_mod
    rng = np.random.default_rng(0)

    def cloud(n, shift):
        p = rng.uniform(0, 1, (n, 3)).astype(np.float32) + np.float32(shift)
        f = rng.uniform(0, 255, (n, 5)).astype(np.float32)
        return p, f

    c = cvo_mod.Cvo(tum_calib, api=oracle_plain)
    api, h = c.api, c.h
    assert c.match_odometry(np.zeros((64, 64, 3), np.uint8), np.zeros((64, 64), np.uint16)) is None  # not init
    api.set_cloud(h, 0, *cloud(40, 0))
    api.set_cloud(h, 1, *cloud(41, 0))
    c.init = True
    live = lambda: (api.slot_size(h, 0), api.slot_size(h, 1))   # noqa: E731  (the device clouds themselves)
    assert c.get_fixed_and_moving_number() == (40, 41)
    c.update_fixed_pcd()                          # fixed <- moving, moving empty
    assert live() == (41, -1)
    # the getter returns what set_pcd cached, also after the move — like the reference (cvo.cpp:370-371, 578-582)
    assert c.get_fixed_and_moving_number() == (40, 41)
    api.set_cloud(h, 1, *cloud(42, 0))
    # reset_keyframe before any update_previous_pcd: fixed <- moving (cvo.cpp:593-596)
    odo = np.eye(4, dtype=np.float32)
    odo[:3, 3] = [0.1, 0.0, 0.0]
    c.reset_keyframe(odo)
    assert live() == (42, -1)
    assert np.array_equal(c.transform, odo)
    api.set_cloud(h, 1, *cloud(43, 0))
    c.update_previous_pcd()                       # previous <- moving
    assert api.slot_size(h, 2) == 43 and c.pre_pc_init
    api.set_cloud(h, 1, *cloud(44, 0))
    c.reset_keyframe(odo)                         # fixed <- previous, previous <- moving
    assert live() == (43, -1) and api.slot_size(h, 2) == 44
    # reset_initial: R,T = inv(transform * odom); returns its inverse (cvo.cpp:611-618)
    c.transform = odo.copy()
    back = c.reset_initial(odo)
    R, T = api.get_RT(h)
    M = np.eye(4, dtype=np.float32)
    M[:3, :3], M[:3, 3] = R, T
    assert np.allclose(M, np.linalg.inv(odo @ odo), atol=1e-6)
    assert np.allclose(back, odo @ odo, atol=1e-6)
    c.close()


def test_track_sequence_on_oracle(oracle_api, tum_calib):
    """The LocalTracker call pattern (two cvo objects, persistent R/T/ell) recovers a smooth
    synthetic trajectory: keyframe tracking of frame k vs frame 0."""
    from cvo_slam_b200 import cvo as cvo_mod, synth
    scene = synth.make_scene(2)
    poses = synth.trajectory(5, 2)
    frames = [synth.to_numpy(*synth.render(scene, P, tum_calib, 640, 480, noise_seed=20 + k))
              for k, P in enumerate(poses)]
    out = cvo_mod.track_sequence(frames, tum_calib, api=oracle_api)
    assert len(out) == 4
    for k, o in enumerate(out):
        gt_odo = synth.relative_transform(poses[k], poses[k + 1])
        gt_kf = synth.relative_transform(poses[0], poses[k + 1])
        ang, dist = pose_error(o["odometry"], gt_odo)
        assert ang < 6e-3 and dist < 6e-3, ("odometry", k, ang, dist)
        ang, dist = pose_error(o["keyframe"], gt_kf)
        assert ang < 8e-3 and dist < 8e-3, ("keyframe", k, ang, dist)
        assert 0 < o["r_odometry"]["cos_angle"] <= 1.01


def _build_dropin(tmp_path):
    from cvo_slam_b200 import build, capi
    build.build()
    exe = os.path.join(str(tmp_path), "dropin_smoke")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "dropin_smoke.cpp"), "-o", exe, "-L", libdir, "-lcvo_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_cpp_dropin_header_compiles_and_links(tmp_path):
    """include/cvo.hpp (the drop-in `cvo::cvo`) compiles against the C ABI and links to the library."""
    exe = _build_dropin(tmp_path)
    assert os.path.exists(exe)
    src = open(os.path.join(ROOT, "include", "cvo.hpp")).read()
    for name in ("set_pcd", "match_odometry", "match_keyframe", "align", "function_inner_product",
                 "compute_innerproduct", "compute_innerproduct_lc", "se3_Hessian", "update_fixed_pcd",
                 "update_previous_pcd", "reset_keyframe", "reset_transform", "reset_initial",
                 "get_fixed_and_moving_number", "get_iteration_number", "get_A_nonzero",
                 "get_fixed_frame_selected_points", "get_moving_frame_selected_points",
                 "first_frame", "prev_transform", "accum_transform"):
        assert name in src, name


def test_reset_initial_arithmetic_header_equals_python_mirror(tmp_path):
    """cvo::reset_initial (cvo.cpp:611-618) is host arithmetic: (transform * odom).inverse().  The drop-in
    header (detail::mul44, detail::inv_affine: Eigen's cofactor inverse restated in float) and the Python
    mirror (_mul44_f32, _inv_affine_f32) must agree to the bit — the alignment that starts from this
    prior amplifies last-bit differences — and both must be an inverse to float accuracy."""
    from cvo_slam_b200 import cvo as cvo_mod, synth
    exe = os.path.join(str(tmp_path), "host_math")
    # the header only needs cvo_b200.h for declarations here: nothing from the library is called
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_math.cpp"), "-o", exe,
                    "-Wl,--unresolved-symbols=ignore-all"], check=True)
    rng = np.random.default_rng(12)
    cases = []
    for _ in range(64):
        A = synth.pose(rng.normal(0, 0.3, 3), rng.normal(0, 0.5, 3)).astype(np.float32)
        B = synth.pose(rng.normal(0, 0.05, 3), rng.normal(0, 0.05, 3)).astype(np.float32)
        cases.append((A, B))
    text = "\n".join(" ".join(repr(float(v)) for v in np.concatenate([A.reshape(-1), B.reshape(-1)])) for A, B in cases)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(cases)
    for (A, B), line in zip(cases, out):
        bits = np.array([int(x, 16) for x in line.split()], dtype=np.uint32).view(np.float32)
        C_cpp, I_cpp = bits[:16].reshape(4, 4), bits[16:].reshape(4, 4)
        C_py = cvo_mod._mul44_f32(A, B)
        I_py = cvo_mod._inv_affine_f32(C_py)
        assert np.array_equal(C_cpp.view(np.uint32), C_py.view(np.uint32))
        assert np.array_equal(I_cpp.view(np.uint32), I_py.view(np.uint32))
        assert np.abs(I_py.astype(np.float64) @ C_py.astype(np.float64) - np.eye(4)).max() < 5e-7


def test_c_abi_struct_layouts_match_ctypes_mirrors(tmp_path):
    """The structs that cross the C ABI by value (include/cvo_b200.h) have the size and field offsets the
    ctypes / numpy mirrors in cvo_slam_b200/capi.py assume (a silent drift would corrupt results)."""
    import ctypes as C
    from cvo_slam_b200 import capi
    src = os.path.join(str(tmp_path), "layout.c")
    exe = os.path.join(str(tmp_path), "layout")
    with open(src, "w") as f:
        f.write('#include <stddef.h>\n#include <stdio.h>\n#include "cvo_b200.h"\n'
                'int main(void) {\n'
                '  printf("lc %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cvo_lc_result), offsetof(cvo_lc_result, num),'
                ' offsetof(cvo_lc_result, post_hessian), offsetof(cvo_lc_result, inliers_svd),'
                ' offsetof(cvo_lc_result, inliers_pnpransac), offsetof(cvo_lc_result, cos_angle), offsetof(cvo_lc_result, accept));\n'
                '  printf("res %zu %zu %zu %zu %zu %zu\\n", sizeof(cvo_align_result), offsetof(cvo_align_result, R),'
                ' offsetof(cvo_align_result, T), offsetof(cvo_align_result, ell), offsetof(cvo_align_result, iterations),'
                ' offsetof(cvo_align_result, last_iter_transform));\n'
                '  printf("pair %zu %zu %zu %zu\\n", sizeof(cvo_pair_desc), offsetof(cvo_pair_desc, R), offsetof(cvo_pair_desc, T),'
                ' offsetof(cvo_pair_desc, ell));\n'
                '  return 0;\n}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    out = dict((l.split()[0], [int(x) for x in l.split()[1:]]) for l in
               subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    L = capi.LcResult
    assert out["lc"] == [C.sizeof(L), L.num.offset, L.post_hessian.offset, L.inliers_svd.offset,
                         L.inliers_pnpransac.offset, L.cos_angle.offset, L.accept.offset]
    assert capi.LC_DTYPE.itemsize == out["lc"][0]
    assert [capi.LC_DTYPE.fields[k][1] for k in ("num", "post_hessian", "inliers_svd", "inliers_pnpransac", "cos_angle", "accept")] == out["lc"][1:]
    R = capi.AlignResult
    assert out["res"] == [C.sizeof(R), R.R.offset, R.T.offset, R.ell.offset, R.iterations.offset,
                          R.last_iter_transform.offset]
    assert capi.RESULT_DTYPE.fields["last_iter_transform"][1] == R.last_iter_transform.offset
    assert capi.RESULT_DTYPE.itemsize == out["res"][0]
    P = capi.PairDesc
    assert out["pair"] == [C.sizeof(P), P.R.offset, P.T.offset, P.ell.offset]
    assert capi.PAIR_DTYPE.itemsize == out["pair"][0]


def test_prev_and_accum_transform_follow_the_last_iteration(oracle_api):
    """cvo.cpp:815-816: prev_transform / accum_transform take `transform` as update_tf() left it at the top
    of the LAST executed iteration (the final update_tf() comes after).  Checked on the oracle: that matrix is
    the final transform of the same alignment stopped one iteration earlier; and the Python mirror of the
    class accumulates it (not the previous call's result)."""
    from cvo_slam_b200 import capi, cvo as cvo_mod, synth
    cal = capi.TUM1_CALIB()
    a, da, b, db, _ = synth.make_pair(7, cal, rot_deg=0.8, trans=(0.015, -0.01, 0.012))
    c = cvo_mod.Cvo(cal, api=oracle_api)
    c.set_pcd(a, da)
    c.match_odometry(b, db)
    full = c.last_result
    n = full.iterations
    assert n > 3
    p = oracle_api.default_params()
    p.max_iter = n - 1
    c2 = cvo_mod.Cvo(cal, p, api=oracle_api)
    c2.set_pcd(a, da)
    c2.match_odometry(b, db)
    assert c2.last_result.iterations == n - 1
    assert np.array_equal(full.last_iter_transform_np(), c2.last_result.transform_np())
    assert not np.array_equal(full.last_iter_transform_np(), full.transform_np())
    assert np.array_equal(c.prev_transform, full.last_iter_transform_np())
    assert np.array_equal(c.accum_transform, full.last_iter_transform_np())   # identity * last
    # a second align of the same object multiplies the new last-iteration transform in
    c.set_pcd(b, db)
    c.align()
    want = cvo_mod._mul44_f32(full.last_iter_transform_np(), c.last_result.last_iter_transform_np())
    assert np.array_equal(c.accum_transform, want)
    c.close()
    c2.close()


def test_bench_roofline_sources_are_committed():
    """bench.py quotes the DRAM traffic of its dominant kernel from committed `ncu --set full` summaries of the very
    workload it runs by default, and from nowhere else: the files exist, parse, name the kernel, and any other
    configuration gets no traffic figure (a number taken under the profiler is never scaled to another size)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for mode, kernel in ((0, "k_align_batch<1>"), (1, "k_align_batch<0>")):
        traffic, src = bench.committed_ncu_traffic(1024, 8, mode)
        assert traffic and traffic > 1e10, (mode, traffic)
        path = os.path.join(ROOT, src)
        assert os.path.exists(path) and kernel in open(path).read()
    assert bench.committed_ncu_traffic(128, 8, 0) == (None, None)
    assert bench.committed_ncu_traffic(1024, 4, 1) == (None, None)
    # the partition of one list of pairs over the ranks covers every pair exactly once
    from cvo_slam_b200 import parallel
    for n, world in ((8192, 8), (8192, 3), (10, 4), (3, 8)):
        parts = [parallel.partition_blocks(n, r, world) for r in range(world)]
        assert sorted(int(i) for p in parts for i in p) == list(range(n))
