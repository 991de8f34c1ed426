"""ctypes binding of the C ABI declared in include/cvo_b200.h.

`load()` returns the binding of libcvo_b200.so (the CUDA product).  It raises if the
shared library is missing: there is no CPU fallback in the product path.

`LowLevel` is written against a symbol prefix so that the test-only CPU oracle
(prefix ``oracle_``, loaded only by the tests) can be driven through the very same Python
surface; the product never imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVO_B200_LIB") or os.path.join(HERE, "libcvo_b200.so")   # override: kernel-variant experiments only

SLOT_FIXED, SLOT_MOVING, SLOT_PREVIOUS = 0, 1, 2

CVO_OK = 0
ERRORS = {
    -1: "CVO_ERR_INVALID",
    -2: "CVO_ERR_CUDA",
    -3: "CVO_ERR_NOT_INIT",
    -4: "CVO_ERR_CAPACITY",
    -5: "CVO_ERR_PAIR_OVERFLOW",
}


class CvoError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: {ERRORS.get(code, code)} {detail}".strip())


class Calib(C.Structure):
    """thirdparty/cvo/include/data_type.h:32-38"""
    _fields_ = [("scaling_factor", C.c_float), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float)]


class Params(C.Structure):
    """SURVEY §9 / cvo.cpp:35-51"""
    _fields_ = [("ell_init", C.c_float), ("sigma", C.c_float), ("sp_thres", C.c_float),
                ("c", C.c_float), ("d", C.c_float), ("c_ell", C.c_float), ("c_sigma", C.c_float),
                ("max_iter", C.c_int32), ("min_step", C.c_float), ("max_step", C.c_float),
                ("eps", C.c_float), ("eps_2", C.c_float), ("ell_after_k2", C.c_float),
                ("ell_after_k9", C.c_float), ("ell_after_k19", C.c_float),
                ("num_want", C.c_int32), ("feature_type", C.c_int32), ("gray_mode", C.c_int32),
                ("exp_mode", C.c_int32)]


class AlignResult(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("R", C.c_float * 9), ("T", C.c_float * 3),
                ("ell", C.c_float), ("iterations", C.c_int32), ("iter", C.c_int32),
                ("A_nonzero", C.c_int32), ("status", C.c_int32), ("num_fixed", C.c_int32), ("num_moving", C.c_int32),
                ("last_iter_transform", C.c_float * 16)]

    def last_iter_transform_np(self):
        return np.array(self.last_iter_transform, dtype=np.float32).reshape(4, 4)

    def transform_np(self):
        return np.array(self.transform, dtype=np.float32).reshape(4, 4)

    def R_np(self):
        return np.array(self.R, dtype=np.float32).reshape(3, 3)

    def T_np(self):
        return np.array(self.T, dtype=np.float32)


class IterRecord(C.Structure):
    _fields_ = [("ell", C.c_float), ("omega", C.c_float * 3), ("v", C.c_float * 3),
                ("B", C.c_double), ("C", C.c_double), ("D", C.c_double), ("E", C.c_double),
                ("step", C.c_float), ("nnz", C.c_int32)]

    def as_dict(self):
        return dict(ell=self.ell, omega=np.array(self.omega, dtype=np.float32),
                    v=np.array(self.v, dtype=np.float32), B=self.B, C=self.C, D=self.D, E=self.E,
                    step=self.step, nnz=self.nnz)


class PairDesc(C.Structure):
    _fields_ = [("fixed_frame", C.c_int32), ("moving_frame", C.c_int32), ("R", C.c_float * 9),
                ("T", C.c_float * 3), ("ell", C.c_float)]


class LcResult(C.Structure):
    """cvo_lc_result (include/cvo_b200.h): compute_innerproduct_lc outputs + the caller's accept rule"""
    _fields_ = [("value", C.c_float * 6), ("num", C.c_int32 * 6), ("post_hessian", C.c_double * 36),
                ("inliers_svd", C.c_int32), ("inliers_pnpransac", C.c_int32), ("cos_angle", C.c_float),
                ("accept", C.c_int32)]


LC_NAMES = ("inn_prior", "inn_lc_prior", "inn_lc_pre", "inn_lc_post", "inn_fixed_pcd", "inn_moving_pcd")
LC_DTYPE = np.dtype([("value", "<f4", (6,)), ("num", "<i4", (6,)), ("post_hessian", "<f8", (36,)),
                     ("inliers_svd", "<i4"), ("inliers_pnpransac", "<i4"), ("cos_angle", "<f4"),
                     ("accept", "<i4")])
assert LC_DTYPE.itemsize == C.sizeof(LcResult)
PAIR_DTYPE = np.dtype([("fixed_frame", "<i4"), ("moving_frame", "<i4"), ("R", "<f4", (9,)),
                       ("T", "<f4", (3,)), ("ell", "<f4")])
RESULT_DTYPE = np.dtype([("transform", "<f4", (16,)), ("R", "<f4", (9,)), ("T", "<f4", (3,)),
                         ("ell", "<f4"), ("iterations", "<i4"), ("iter", "<i4"),
                         ("A_nonzero", "<i4"), ("status", "<i4"), ("num_fixed", "<i4"), ("num_moving", "<i4"),
                         ("last_iter_transform", "<f4", (16,))])
assert PAIR_DTYPE.itemsize == C.sizeof(PairDesc)
assert RESULT_DTYPE.itemsize == C.sizeof(AlignResult)


def TUM1_CALIB():
    """config/TUM1.yaml:8-20"""
    return Calib(5000.0, 517.306408, 516.469215, 318.643040, 255.313989)


def ETH3D_CALIB():
    """config/ETH3D_training_1.yaml:10-22"""
    return Calib(5000.0, 726.28741455078, 726.28741455078, 354.6496887207, 186.46566772461)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class LowLevel:
    """Python surface of the handle API (include/cvo_b200.h), for a given symbol prefix."""

    def __init__(self, lib, prefix="cvo_"):
        self.lib = lib
        self.prefix = prefix
        self._declare()

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _declare(self):
        P = C.POINTER
        vp = C.c_void_p
        sig = {
            "default_params": ([P(Params)], None),
            "destroy": ([vp], C.c_int),
            "set_frame": ([vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.c_int], C.c_int),
            "set_cloud": ([vp, C.c_int, C.c_int, P(C.c_float), P(C.c_float)], C.c_int),
            "slot_move": ([vp, C.c_int, C.c_int], C.c_int),
            "slot_size": ([vp, C.c_int, P(C.c_int)], C.c_int),
            "set_RT": ([vp, P(C.c_float), P(C.c_float)], C.c_int),
            "get_RT": ([vp, P(C.c_float), P(C.c_float)], C.c_int),
            "set_ell": ([vp, C.c_float], C.c_int),
            "get_ell": ([vp, P(C.c_float)], C.c_int),
            "align": ([vp, P(AlignResult), P(IterRecord), C.c_int], C.c_int),
            "iteration_at": ([vp, P(C.c_float), P(C.c_float), C.c_float, P(IterRecord)], C.c_int),
            "last_pattern": ([vp, P(C.c_int32), P(C.c_float), C.c_int, P(C.c_int)], C.c_int),
            "inner_product": ([vp, C.c_int, P(C.c_float), C.c_int, P(C.c_float), P(C.c_int)], C.c_int),
            "hessian": ([vp, C.c_int, P(C.c_float), C.c_int, P(C.c_double), P(C.c_int)], C.c_int),
            "get_selected_points": ([vp, C.c_int, P(C.c_float), C.c_int, P(C.c_int)], C.c_int),
            "get_cloud": ([vp, C.c_int, P(C.c_float), P(C.c_float), C.c_int, P(C.c_int)], C.c_int),
            "get_selection_debug": ([vp, C.c_int, vp, P(C.c_int32)], C.c_int),
        }
        for name, (args, res) in sig.items():
            f = self._fn(name)
            f.argtypes = args
            f.restype = res

    def _check(self, rc, where):
        if rc != CVO_OK:
            detail = ""
            if self.prefix == "cvo_":
                self.lib.cvo_last_error.restype = C.c_char_p
                detail = (self.lib.cvo_last_error() or b"").decode()
            raise CvoError(rc, self.prefix + where, detail)

    # -- construction ------------------------------------------------------------------------
    def default_params(self):
        p = Params()
        self._fn("default_params")(C.byref(p))
        return p

    def create(self, calib, params=None, device=0):
        raise NotImplementedError

    def destroy(self, h):
        self._fn("destroy")(h)

    # -- clouds ------------------------------------------------------------------------------
    def set_frame(self, h, slot, bgr, depth):
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        hh, ww = depth.shape
        assert bgr.shape == (hh, ww, 3)
        self._check(self._fn("set_frame")(h, slot, bgr.ctypes.data, 3 * ww, depth.ctypes.data,
                                          2 * ww, ww, hh), "set_frame")

    def set_cloud(self, h, slot, pos, feat):
        pos = _f32(pos, (-1, 3))
        feat = _f32(feat, (-1, 5))
        assert len(pos) == len(feat)
        self._check(self._fn("set_cloud")(h, slot, len(pos), _fp(pos), _fp(feat)), "set_cloud")

    def slot_move(self, h, dst, src):
        self._check(self._fn("slot_move")(h, dst, src), "slot_move")

    def slot_size(self, h, slot):
        n = C.c_int(0)
        rc = self._fn("slot_size")(h, slot, C.byref(n))
        return n.value if rc == CVO_OK else -1

    def get_cloud(self, h, slot):
        n = self.slot_size(h, slot)
        if n < 0:
            raise CvoError(-3, self.prefix + "get_cloud")
        pos = np.zeros((max(n, 1), 3), np.float32)
        feat = np.zeros((max(n, 1), 5), np.float32)
        m = C.c_int(0)
        self._check(self._fn("get_cloud")(h, slot, _fp(pos), _fp(feat), n, C.byref(m)), "get_cloud")
        return pos[:n], feat[:n]

    def get_selected_points(self, h, slot):
        n = self.slot_size(h, slot)
        if n < 0:
            raise CvoError(-3, self.prefix + "get_selected_points")
        xy = np.zeros((max(n, 1), 2), np.float32)
        m = C.c_int(0)
        self._check(self._fn("get_selected_points")(h, slot, _fp(xy), n, C.byref(m)),
                    "get_selected_points")
        return xy[:n]

    def get_selection_debug(self, h, slot, w, hgt):
        m = np.zeros((hgt, w), np.uint8)
        info = (C.c_int32 * 5)()
        self._check(self._fn("get_selection_debug")(h, slot, m.ctypes.data, info),
                    "get_selection_debug")
        return m, dict(n2=info[0], n3=info[1], n4=info[2], pot=info[3], passes=info[4])

    # -- state -------------------------------------------------------------------------------
    def set_RT(self, h, R, T):
        R = _f32(R, (9,))
        T = _f32(T, (3,))
        self._check(self._fn("set_RT")(h, _fp(R), _fp(T)), "set_RT")

    def get_RT(self, h):
        R = np.zeros(9, np.float32)
        T = np.zeros(3, np.float32)
        self._check(self._fn("get_RT")(h, _fp(R), _fp(T)), "get_RT")
        return R.reshape(3, 3), T

    def set_ell(self, h, ell):
        self._check(self._fn("set_ell")(h, float(ell)), "set_ell")

    def get_ell(self, h):
        e = C.c_float(0)
        self._check(self._fn("get_ell")(h, C.byref(e)), "get_ell")
        return e.value

    # -- the hot path ------------------------------------------------------------------------
    def align(self, h, trace_cap=0):
        out = AlignResult()
        trace = (IterRecord * max(trace_cap, 1))()
        self._check(self._fn("align")(h, C.byref(out), trace if trace_cap else None, trace_cap),
                    "align")
        recs = [trace[i].as_dict() for i in range(min(trace_cap, out.iterations))]
        return out, recs

    def iteration_at(self, h, R, T, ell):
        R = _f32(R, (9,))
        T = _f32(T, (3,))
        rec = IterRecord()
        self._check(self._fn("iteration_at")(h, _fp(R), _fp(T), float(ell), C.byref(rec)),
                    "iteration_at")
        return rec.as_dict()

    def last_pattern(self, h, cap):
        ij = np.zeros((max(cap, 1), 2), np.int32)
        a = np.zeros(max(cap, 1), np.float32)
        n = C.c_int(0)
        self._check(self._fn("last_pattern")(h, ij.ctypes.data_as(C.POINTER(C.c_int32)), _fp(a), cap,
                                             C.byref(n)), "last_pattern")
        m = min(n.value, cap)
        return ij[:m], a[:m], n.value

    def inner_product(self, h, slot_a, Ta, slot_b):
        val = C.c_float(0)
        num = C.c_int(0)
        ta = None if Ta is None else _f32(np.asarray(Ta)[:3, :4], (12,))
        self._check(self._fn("inner_product")(h, slot_a, None if ta is None else _fp(ta), slot_b,
                                              C.byref(val), C.byref(num)), "inner_product")
        return val.value, num.value

    def hessian(self, h, slot_a, Ta, slot_b):
        H = np.zeros(36, np.float64)
        inl = C.c_int(0)
        ta = None if Ta is None else _f32(np.asarray(Ta)[:3, :4], (12,))
        self._check(self._fn("hessian")(h, slot_a, None if ta is None else _fp(ta), slot_b,
                                        H.ctypes.data_as(C.POINTER(C.c_double)), C.byref(inl)),
                    "hessian")
        return H.reshape(6, 6), inl.value


class CudaLowLevel(LowLevel):
    """libcvo_b200.so"""

    def __init__(self, lib):
        super().__init__(lib, "cvo_")
        P = C.POINTER
        vp = C.c_void_p
        lib.cvo_create.argtypes = [P(Calib), P(Params), C.c_int, P(vp)]
        lib.cvo_create.restype = C.c_int
        lib.cvo_last_error.restype = C.c_char_p
        lib.cvo_random_pattern.argtypes = [vp, C.c_int]
        lib.cvo_set_frame_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int]
        lib.cvo_handle_stats.argtypes = [vp, P(C.c_int64)]
        lib.cvo_batch_create.argtypes = [P(Calib), P(Params), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, P(vp)]
        lib.cvo_batch_destroy.argtypes = [vp]
        lib.cvo_batch_set_frames.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        lib.cvo_batch_set_frames_device.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        lib.cvo_batch_frame_size.argtypes = [vp, C.c_int, P(C.c_int)]
        lib.cvo_batch_align.argtypes = [vp, C.c_int, vp, vp]
        lib.cvo_batch_inner_product.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        lib.cvo_batch_stats.argtypes = [vp, P(C.c_int64)]
        lib.cvo_batch_verify_lc.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
        lib.cvo_compute_innerproduct_lc.argtypes = [vp, vp, vp, vp, vp, P(LcResult)]
        lib.cvo_batch_last_align_ms.argtypes = [vp, P(C.c_float)]
        lib.cvo_batch_mark.argtypes = [vp, C.c_int]
        lib.cvo_batch_elapsed_ms.argtypes = [vp, P(C.c_float)]
        lib.cvo_multi_create.argtypes = [P(Calib), P(Params), C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, P(vp)]
        lib.cvo_multi_destroy.argtypes = [vp]
        lib.cvo_multi_align.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
        lib.cvo_multi_last_shares.argtypes = [vp, vp, vp, vp]

    def create(self, calib, params=None, device=0):
        if params is None:
            params = self.default_params()
        h = C.c_void_p()
        self._check(self.lib.cvo_create(C.byref(calib), C.byref(params), device, C.byref(h)),
                    "create")
        return h

    def copy_cloud(self, dst, dst_slot, src, src_slot):
        self.lib.cvo_copy_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self._check(self.lib.cvo_copy_cloud(dst, dst_slot, src, src_slot), "copy_cloud")

    def random_pattern(self, n):
        out = np.zeros(n, np.uint8)
        self._check(self.lib.cvo_random_pattern(out.ctypes.data, n), "random_pattern")
        return out

    def set_frame_device(self, h, slot, bgr_ptr, depth_ptr, w, hgt):
        self._check(self.lib.cvo_set_frame_device(h, slot, bgr_ptr, depth_ptr, w, hgt),
                    "set_frame_device")

    def phase_cycles(self, h):
        s = (C.c_int64 * 8)()
        self.lib.cvo_handle_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        self._check(self.lib.cvo_handle_phase_cycles(h, s), "handle_phase_cycles")
        return dict(zip(("grid", "P0", "P1a", "P1b", "P2", "P3", "filters", "rebuilds"), [int(x) for x in s]))

    def compute_innerproduct(self, h, tran):
        """cvo_compute_innerproduct: ((value, num) x 4, H, inliers) in one launch"""
        self.lib.cvo_compute_innerproduct.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                      C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        T = _f32(tran, (16,))
        v = np.zeros(4, np.float32)
        n = np.zeros(4, np.int32)
        H = np.zeros(36, np.float64)
        inl = C.c_int(0)
        self._check(self.lib.cvo_compute_innerproduct(h, _fp(T), _fp(v), n.ctypes.data_as(C.POINTER(C.c_int)),
                                                      H.ctypes.data_as(C.POINTER(C.c_double)), C.byref(inl)),
                    "compute_innerproduct")
        return [(float(v[k]), int(n[k])) for k in range(4)], H.reshape(6, 6), inl.value

    def compute_innerproduct_lc(self, h, prior_tran, lc_prior_tran, lc_prior_tran_2, lc_tran):
        """cvo_compute_innerproduct_lc: the eight queries of cvo.cpp:505-561 in one launch -> LcResult"""
        Ts = [_f32(m, (16,)) for m in (prior_tran, lc_prior_tran, lc_prior_tran_2, lc_tran)]
        out = LcResult()
        self._check(self.lib.cvo_compute_innerproduct_lc(h, Ts[0].ctypes.data, Ts[1].ctypes.data, Ts[2].ctypes.data,
                                                         Ts[3].ctypes.data, C.byref(out)), "compute_innerproduct_lc")
        return out

    def get_selected_points_device(self, h, slot):
        """-> (device pointer as int, n): the selected pixels as n (x, y) float pairs on the device"""
        ptr = C.c_void_p()
        n = C.c_int(0)
        self.lib.cvo_get_selected_points_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        self._check(self.lib.cvo_get_selected_points_device(h, slot, C.byref(ptr), C.byref(n)), "get_selected_points_device")
        return ptr.value, n.value

    def handle_stats(self, h):
        s = (C.c_int64 * 4)()
        self._check(self.lib.cvo_handle_stats(h, s), "handle_stats")
        return dict(launches=s[0], evals=s[1], iterations=s[2], nnz=s[3])


_lib = None


def load():
    """Load libcvo_b200.so.  Fails loudly if it has not been built (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  The product has no CPU fallback.")
        _lib = CudaLowLevel(C.CDLL(LIB_PATH))
    return _lib
