"""ctypes loader of the CPU oracle (oracle/cvo_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under cvo_slam_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from cvo_slam_b200.capi import (Calib, LowLevel, Params, CVO_OK)  # noqa: E402  (struct layouts only)

LIB = os.path.join(HERE, "libcvo_oracle.so")
LIB_KD = os.path.join(HERE, "_ref", "libcvo_oracle_kd.so")
LIB_REFSEL = os.path.join(HERE, "_ref", "libref_select.so")   # the reference's own selector sources, compiled here
LIB_REFCVO = os.path.join(HERE, "_ref", "libref_cvo.so")      # the reference's own cvo.cpp / LieGroup.cpp, compiled here

SEARCH_GRID, SEARCH_BRUTE, SEARCH_NANOFLANN = 0, 1, 2


def build(force=False):
    """Compile the oracle (and, where /root/reference exists, the nanoflann variant)."""
    src = os.path.join(HERE, "cvo_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/thirdparty/cvo/thirdparty/nanoflann.hpp"):
        if force or not os.path.exists(LIB_KD) or os.path.getmtime(LIB_KD) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/thirdparty/cvo/src/pcd_generator.cpp"):
        subprocess.check_call(["make", "-C", HERE, "refsel", "refcvo"], stdout=subprocess.DEVNULL)


class OracleLowLevel(LowLevel):
    def __init__(self, lib):
        super().__init__(lib, "oracle_")
        P = C.POINTER
        vp = C.c_void_p
        lib.oracle_create.argtypes = [P(Calib), P(Params), P(vp)]
        lib.oracle_set_search.argtypes = [vp, C.c_int]
        lib.oracle_stats.argtypes = [vp, P(C.c_int64)]
        lib.oracle_gray.argtypes = [vp, C.c_int, C.c_int, vp]
        lib.oracle_hsv.argtypes = [vp, C.c_int, vp]
        lib.oracle_stages.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
        lib.oracle_random_pattern.argtypes = [vp, C.c_int]
        lib.oracle_cubic_real_roots.argtypes = [C.c_double] * 4 + [P(C.c_double)]
        lib.oracle_step_from_coeffs.argtypes = [C.c_double] * 4 + [C.c_float, C.c_float]
        lib.oracle_step_from_coeffs.restype = C.c_float
        lib.oracle_exp_sek3.argtypes = [P(C.c_float), P(C.c_float), C.c_float, P(C.c_float), P(C.c_float)]
        lib.oracle_dist_se3.argtypes = [P(C.c_float), P(C.c_float)]
        lib.oracle_dist_se3.restype = C.c_float
        lib.oracle_finish_hessian.argtypes = [P(C.c_float), C.c_int, P(C.c_double)]
        lib.oracle_radius_search.argtypes = [P(C.c_float), C.c_int, P(C.c_float), C.c_float, C.c_int,
                                             P(C.c_int32), P(C.c_float), C.c_int]
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        self.has_nanoflann = bool(lib.oracle_has_nanoflann())
        self.search_mode = SEARCH_NANOFLANN if self.has_nanoflann else SEARCH_GRID

    def num_threads(self):
        return int(self.lib.oracle_num_threads())

    def set_num_threads(self, n):
        self.lib.oracle_set_num_threads(int(n))

    def create(self, calib, params=None, device=0, search=None):
        if params is None:
            params = self.default_params()
        h = C.c_void_p()
        self._check(self.lib.oracle_create(C.byref(calib), C.byref(params), C.byref(h)), "create")
        self.lib.oracle_set_search(h, self.search_mode if search is None else search)
        return h

    def set_search(self, h, mode):
        self.lib.oracle_set_search(h, mode)

    def stats(self, h):
        s = (C.c_int64 * 3)()
        self.lib.oracle_stats(h, s)
        return dict(launches=0, evals=s[1], iterations=s[2])

    # ---- stage-level helpers -------------------------------------------------------------
    def gray(self, bgr, mode=0):
        bgr = np.ascontiguousarray(bgr, np.uint8).reshape(-1, 3)
        out = np.zeros(len(bgr), np.uint8)
        self.lib.oracle_gray(bgr.ctypes.data, len(bgr), mode, out.ctypes.data)
        return out

    def hsv(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8).reshape(-1, 3)
        out = np.zeros((len(bgr), 3), np.uint8)
        self.lib.oracle_hsv(bgr.ctypes.data, len(bgr), out.ctypes.data)
        return out

    def stages(self, bgr, gray_mode=0):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        sizes = [w * h, (w // 2) * (h // 2), (w // 4) * (h // 4)]
        gray = np.zeros((h, w), np.uint8)
        g2 = np.zeros(sum(sizes), np.float32)
        nb = (w // 32) * (h // 32)
        ths = np.zeros(nb, np.float32)
        thss = np.zeros(nb, np.float32)
        dx0 = np.zeros((h, w), np.float32)
        dy0 = np.zeros((h, w), np.float32)
        self.lib.oracle_stages(bgr.ctypes.data, w, h, gray_mode, gray.ctypes.data, g2.ctypes.data,
                               ths.ctypes.data, thss.ctypes.data, dx0.ctypes.data, dy0.ctypes.data)
        lv = [g2[:sizes[0]].reshape(h, w),
              g2[sizes[0]:sizes[0] + sizes[1]].reshape(h // 2, w // 2),
              g2[sizes[0] + sizes[1]:].reshape(h // 4, w // 4)]
        return dict(gray=gray, g2=lv, ths=ths.reshape(h // 32, w // 32),
                    ths_smoothed=thss.reshape(h // 32, w // 32), dx0=dx0, dy0=dy0)

    def random_pattern(self, n):
        out = np.zeros(n, np.uint8)
        self.lib.oracle_random_pattern(out.ctypes.data, n)
        return out

    def cubic_real_roots(self, a, b, c, d):
        re = (C.c_double * 3)()
        n = self.lib.oracle_cubic_real_roots(a, b, c, d, re)
        return [re[i] for i in range(n)]

    def step_from_coeffs(self, B, Cc, D, E, min_step=0.2, max_step=0.8):
        return float(self.lib.oracle_step_from_coeffs(B, Cc, D, E, min_step, max_step))

    def exp_sek3(self, w, v, dt):
        w = np.ascontiguousarray(w, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        R = np.zeros(9, np.float32)
        t = np.zeros(3, np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        self.lib.oracle_exp_sek3(fp(w), fp(v), float(dt), fp(R), fp(t))
        return R.reshape(3, 3), t

    def dist_se3(self, R, T):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        T = np.ascontiguousarray(T, np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        return float(self.lib.oracle_dist_se3(fp(R), fp(T)))

    def finish_hessian(self, Hf, inliers):
        Hf = np.ascontiguousarray(Hf, np.float32).reshape(36)
        out = np.zeros(36, np.float64)
        self.lib.oracle_finish_hessian(Hf.ctypes.data_as(C.POINTER(C.c_float)), int(inliers),
                                       out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.reshape(6, 6)

    def radius_search(self, pts, q, radius2, mode):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        q = np.ascontiguousarray(q, np.float32)
        cap = len(pts)
        idx = np.zeros(cap, np.int32)
        d2 = np.zeros(cap, np.float32)
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
        n = self.lib.oracle_radius_search(fp(pts), len(pts), fp(q), float(radius2), mode,
                                          idx.ctypes.data_as(C.POINTER(C.c_int32)), fp(d2), cap)
        return idx[:n], d2[:n]


_cache = {}


def load(kd=None):
    """kd=True: the oracle/_ref build whose radius search is the reference's nanoflann;
    kd=False: the self-contained build; kd=None: nanoflann when available."""
    build()
    if kd is None:
        kd = os.path.exists(LIB_KD)
    path = LIB_KD if kd else LIB
    if path not in _cache:
        if not os.path.exists(path):
            raise ImportError(f"{path} not built")
        _cache[path] = OracleLowLevel(C.CDLL(path))
    return _cache[path]


class RefSelect:
    """The reference's own pcd_generator::create_pointcloud (pcd_generator.cpp + PixelSelector2.cpp compiled
    where they lie, oracle/ref_select.cpp): status map, selected pixels, positions, features of one frame."""

    def __init__(self, lib):
        self.lib = lib
        vp = C.c_void_p
        lib.refsel_run.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp]
        lib.refsel_run.restype = C.c_int

    def run(self, bgr, depth, calib, num_want=3000, feature_type=1, gray_mode=0):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        h, w = depth.shape
        cal = np.array([calib.scaling_factor, calib.fx, calib.fy, calib.cx, calib.cy], np.float32)
        cap = w * h
        smap = np.zeros((h, w), np.float32)
        gray = np.zeros((h, w), np.float32)
        pix = np.zeros((cap, 2), np.float32)
        pos = np.zeros((cap, 3), np.float32)
        feat = np.zeros((cap, 5), np.float32)
        n = self.lib.refsel_run(bgr.ctypes.data, depth.ctypes.data, w, h, cal.ctypes.data, int(num_want), int(feature_type),
                                int(gray_mode), smap.ctypes.data, pix.ctypes.data, pos.ctypes.data, feat.ctypes.data, cap,
                                gray.ctypes.data)
        assert n >= 0
        return dict(map=smap.astype(np.uint8), gray=gray.astype(np.uint8), pix=pix[:n].copy(), pos=pos[:n].copy(),
                    feat=feat[:n].copy(), n=n)


def load_refsel():
    """-> RefSelect, or None where the reference (and hence oracle/_ref/libref_select.so) is not available."""
    build()
    if not os.path.exists(LIB_REFSEL):
        return None
    if LIB_REFSEL not in _cache:
        _cache[LIB_REFSEL] = RefSelect(C.CDLL(LIB_REFSEL))
    return _cache[LIB_REFSEL]


class _RefRecord(C.Structure):
    _fields_ = [("ell", C.c_float), ("omega", C.c_float * 3), ("v", C.c_float * 3), ("step", C.c_float), ("nnz", C.c_int32)]


class RefCvo:
    """The reference's own cvo::cvo (cvo.cpp, LieGroup.cpp, pcd_generator.cpp, PixelSelector2.cpp and its vendored
    nanoflann compiled where they lie, oracle/ref_cvo.cpp).  One instance = one cvo object."""

    def __init__(self, lib, calib):
        self.lib = lib
        vp = C.c_void_p
        lib.refcvo_create.restype = vp
        lib.refcvo_create.argtypes = [vp]
        lib.refcvo_destroy.argtypes = [vp]
        lib.refcvo_set_pcd.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        lib.refcvo_set_clouds.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
        lib.refcvo_sizes.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.refcvo_set_state.argtypes = [vp, vp, vp, C.c_float]
        lib.refcvo_get_state.argtypes = [vp, vp, vp, C.POINTER(C.c_float), vp]
        lib.refcvo_iteration_at.argtypes = [vp, vp, vp, C.c_float, C.POINTER(_RefRecord), C.c_int, vp, vp]
        lib.refcvo_align.argtypes = [vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        lib.refcvo_compute_innerproduct.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        lib.refcvo_compute_innerproduct_lc.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        lib.refcvo_update_fixed_pcd.argtypes = [vp]
        lib.refcvo_reset_initial.argtypes = [vp, vp, vp]
        lib.refcvo_update_previous_pcd.argtypes = [vp]
        lib.refcvo_reset_keyframe.argtypes = [vp, vp]
        lib.refcvo_reset_transform.argtypes = [vp, vp]
        lib.refcvo_slot_sizes.argtypes = [vp, vp]
        lib.refcvo_set_max_iter.argtypes = [vp, C.c_int]
        cal = np.array([calib.scaling_factor, calib.fx, calib.fy, calib.cx, calib.cy], np.float32)
        self.h = lib.refcvo_create(cal.ctypes.data)
        assert self.h

    def close(self):
        if self.h:
            self.lib.refcvo_destroy(self.h)
            self.h = None

    def set_pcd(self, bgr, depth):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        self.lib.refcvo_set_pcd(self.h, bgr.ctypes.data, depth.ctypes.data, depth.shape[1], depth.shape[0])

    def set_clouds(self, pos_f, feat_f, pos_m, feat_m):
        a = [np.ascontiguousarray(x, np.float32) for x in (pos_f, feat_f, pos_m, feat_m)]
        self.lib.refcvo_set_clouds(self.h, len(a[0]), a[0].ctypes.data, a[1].ctypes.data, len(a[2]), a[2].ctypes.data, a[3].ctypes.data)

    def sizes(self):
        nf, nm = C.c_int(0), C.c_int(0)
        self.lib.refcvo_sizes(self.h, C.byref(nf), C.byref(nm))
        return nf.value, nm.value

    def set_state(self, R, T, ell):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        T = np.ascontiguousarray(T, np.float32)
        self.lib.refcvo_set_state(self.h, R.ctypes.data, T.ctypes.data, float(ell))

    def get_state(self):
        R, T, tf, ell = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(16, np.float32), C.c_float(0)
        self.lib.refcvo_get_state(self.h, R.ctypes.data, T.ctypes.data, C.byref(ell), tf.ctypes.data)
        return R.reshape(3, 3), T, ell.value, tf.reshape(4, 4)

    def iteration_at(self, R, T, ell, cap=1 << 21):
        R = np.ascontiguousarray(R, np.float32).reshape(9)
        T = np.ascontiguousarray(T, np.float32)
        rec = _RefRecord()
        ij = np.zeros((cap, 2), np.int32)
        a = np.zeros(cap, np.float32)
        n = self.lib.refcvo_iteration_at(self.h, R.ctypes.data, T.ctypes.data, float(ell), C.byref(rec), cap, ij.ctypes.data, a.ctypes.data)
        return dict(ell=rec.ell, omega=np.array(rec.omega, np.float32), v=np.array(rec.v, np.float32), step=rec.step, nnz=rec.nnz,
                    ij=ij[:n].copy(), a=a[:n].copy())

    def align(self):
        tf, last = np.zeros(16, np.float32), np.zeros(16, np.float32)
        it, nnz, ell = C.c_int(0), C.c_int(0), C.c_float(0)
        self.lib.refcvo_align(self.h, tf.ctypes.data, last.ctypes.data, C.byref(it), C.byref(nnz), C.byref(ell))
        return dict(transform=tf.reshape(4, 4), last_iter_transform=last.reshape(4, 4), iter=it.value, A_nonzero=nnz.value, ell=ell.value)

    def compute_innerproduct(self, tran):
        t = np.ascontiguousarray(tran, np.float32).reshape(16)
        v, n, H = np.zeros(4, np.float32), np.zeros(4, np.int32), np.zeros(36, np.float64)
        inl, cos = C.c_int(0), C.c_float(0)
        self.lib.refcvo_compute_innerproduct(self.h, t.ctypes.data, v.ctypes.data, n.ctypes.data, H.ctypes.data, C.byref(inl), C.byref(cos))
        return dict(values=v, nums=n, H=H.reshape(6, 6), inliers=inl.value, cos_angle=cos.value)

    def compute_innerproduct_lc(self, prior_tran, lc_prior_tran, lc_prior_tran_2, lc_tran):
        """cvo::compute_innerproduct_lc (cvo.cpp:505-561) -> values / nums in the order {inn_prior, inn_lc_prior,
        inn_lc_pre, inn_lc_post, inn_fixed_pcd, inn_moving_pcd}, post_hessian, the two inlier counts, cos_angle"""
        t = [np.ascontiguousarray(x, np.float32).reshape(16) for x in (prior_tran, lc_prior_tran, lc_prior_tran_2, lc_tran)]
        v, n, H = np.zeros(6, np.float32), np.zeros(6, np.int32), np.zeros(36, np.float64)
        i1, i2, cos = C.c_int(0), C.c_int(0), C.c_float(0)
        self.lib.refcvo_compute_innerproduct_lc(self.h, t[0].ctypes.data, t[1].ctypes.data, t[2].ctypes.data, t[3].ctypes.data,
                                                v.ctypes.data, n.ctypes.data, H.ctypes.data, C.byref(i1), C.byref(i2), C.byref(cos))
        return dict(values=v, nums=n, H=H.reshape(6, 6), inliers_svd=i1.value, inliers_pnpransac=i2.value, cos_angle=cos.value)

    def update_fixed_pcd(self):
        self.lib.refcvo_update_fixed_pcd(self.h)

    def reset_initial(self, odom):
        o = np.ascontiguousarray(odom, np.float32).reshape(16)
        back = np.zeros(16, np.float32)
        self.lib.refcvo_reset_initial(self.h, o.ctypes.data, back.ctypes.data)
        return back.reshape(4, 4)

    def update_previous_pcd(self):
        self.lib.refcvo_update_previous_pcd(self.h)

    def reset_keyframe(self, odom):
        o = np.ascontiguousarray(odom, np.float32).reshape(16)
        self.lib.refcvo_reset_keyframe(self.h, o.ctypes.data)

    def reset_transform(self, odom):
        o = np.ascontiguousarray(odom, np.float32).reshape(16)
        self.lib.refcvo_reset_transform(self.h, o.ctypes.data)

    def slot_sizes(self):
        """points in the fixed / moving / previous slot, -1 for an empty (moved-from) slot"""
        n = np.zeros(3, np.int32)
        self.lib.refcvo_slot_sizes(self.h, n.ctypes.data)
        return [int(x) for x in n]

    def set_max_iter(self, n):
        self.lib.refcvo_set_max_iter(self.h, int(n))


def load_refcvo(calib):
    """-> RefCvo (a fresh reference cvo object), or None where the reference is not available."""
    build()
    if not os.path.exists(LIB_REFCVO):
        return None
    if LIB_REFCVO not in _cache:
        _cache[LIB_REFCVO] = C.CDLL(LIB_REFCVO)
    return RefCvo(_cache[LIB_REFCVO], calib)
