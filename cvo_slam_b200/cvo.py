"""Python mirror of the reference's `cvo::cvo` class (thirdparty/cvo/include/cvo.hpp:82-282).

Same method names, argument meaning and error behaviour as the C++ class; the state shuffles
(`update_fixed_pcd`, `reset_keyframe`, `reset_initial`, ...) are host logic, everything
numerical goes through the C ABI (include/cvo_b200.h).  The C++ drop-in with the identical
logic is include/cvo.hpp; this mirror exists so that tests and bench.py read like code
written against the reference class.

`api` defaults to the CUDA library (capi.load(), no CPU fallback).  The test-suite passes the
oracle's LowLevel instead to drive the CPU restatement through the very same host logic.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .capi import SLOT_FIXED, SLOT_MOVING, SLOT_PREVIOUS


def _mul44_f32(a, b):
    """4x4 product in float32, sum over k in order (include/cvo.hpp detail::mul44)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    o = np.zeros((4, 4), np.float32)
    for r in range(4):
        for c in range(4):
            s = np.float32(0)
            for k in range(4):
                s = np.float32(s + np.float32(a[r, k] * b[k, c]))
            o[r, c] = s
    return o


def _inv_affine_f32(a):
    """Eigen::Affine3f::inverse() restated in float32 (include/cvo.hpp detail::inv_affine): cofactor inverse of
    the linear part, translation = -(inverse * t)."""
    a = np.asarray(a, np.float32)
    f = np.float32

    def cof(i, j):
        i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
        return f(f(a[i1, j1] * a[i2, j2]) - f(a[i1, j2] * a[i2, j1]))

    c00, c10, c20 = cof(0, 0), cof(1, 0), cof(2, 0)
    det = f(f(f(c00 * a[0, 0]) + f(c10 * a[1, 0])) + f(c20 * a[2, 0]))
    invdet = f(f(1.0) / det)
    inv = np.zeros((3, 3), np.float32)
    for r in range(3):
        for c in range(3):
            inv[r, c] = f(cof(c, r) * invdet)
    o = np.eye(4, dtype=np.float32)
    o[:3, :3] = inv
    for r in range(3):
        o[r, 3] = -f(f(f(inv[r, 0] * a[0, 3]) + f(inv[r, 1] * a[1, 3])) + f(inv[r, 2] * a[2, 3]))
    return o


class InnP:
    """cvo::inn_p (cvo.hpp:52-80)"""

    def __init__(self, value=0.0, num=0, num_e=0):
        self.value = np.float32(value)
        self.num = int(num)
        self.num_e = int(num_e)

    def __repr__(self):
        return f"inn_p(value={self.value}, num={self.num})"


class Cvo:
    def __init__(self, calib, params=None, api=None, device=0):
        self.api = api if api is not None else capi.load()
        self.h = self.api.create(calib, params, device)
        # public members of the reference class (cvo.hpp:137-146)
        self.first_frame = True
        self.init = False
        self.iter = 0
        self.transform = np.eye(4, dtype=np.float32)
        self.prev_transform = np.eye(4, dtype=np.float32)
        self.accum_transform = np.eye(4, dtype=np.float32)
        # private
        self.pre_pc_init = False
        self.num_fixed = 0
        self.num_moving = 0
        self._sizes_valid = False
        self.A_nonzero = 0
        self.last_result = None

    def close(self):
        if self.h is not None:
            self.api.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- cvo.cpp:345-386 -------------------------------------------------------------------
    def set_pcd(self, rgb_img, dep_img):
        if not self.init:
            self.api.set_frame(self.h, SLOT_FIXED, rgb_img, dep_img)
            self.init = True
            return
        self.api.set_frame(self.h, SLOT_MOVING, rgb_img, dep_img)
        self._sizes_valid = False     # the reference refreshes num_fixed / num_moving here (cvo.cpp:370-371)
        self.A_nonzero = 0

    def set_pcd_from(self, src, src_slot):
        """Opt-in (include/cvo.hpp set_pcd_from): adopt the device cloud `src` has just selected for
        the same image instead of selecting it again."""
        slot = SLOT_MOVING if self.init else SLOT_FIXED
        self.api.copy_cloud(self.h, slot, src.h, src_slot)
        if not self.init:
            self.init = True
            return
        self._sizes_valid = False
        self.A_nonzero = 0

    def match_keyframe_from(self, src, src_slot):
        if not self.init:
            print("cvo not initialized !")
            return None
        self.set_pcd_from(src, src_slot)
        self.align()
        return self.transform.astype(np.float64)

    # ---- cvo.cpp:763-821 -------------------------------------------------------------------
    def align(self, trace_cap=0):
        res, recs = self.api.align(self.h, trace_cap)
        self.last_result = res
        if res.iter >= 0:
            self.iter = res.iter          # `iter` is only written on break (cvo.cpp:783,805)
        self.A_nonzero = res.A_nonzero
        self.num_fixed, self.num_moving, self._sizes_valid = int(res.num_fixed), int(res.num_moving), True
        # cvo.cpp:815-816: `transform` as the last executed iteration's update_tf() left it
        last = res.last_iter_transform_np()
        self.prev_transform = last.copy()
        self.accum_transform = _mul44_f32(self.accum_transform, last)
        self.transform = res.transform_np()
        return recs

    # ---- cvo.cpp:461-473 / 563-576 -----------------------------------------------------------
    def match_odometry(self, rgb_img, dep_img):
        if not self.init:
            print("cvo not initialized !")
            return None
        self.set_pcd(rgb_img, dep_img)
        self.align()
        return self.transform.astype(np.float64)

    def match_keyframe(self, rgb_img, dep_img):
        if not self.init:
            print("cvo not initialized !")
            return None
        self.set_pcd(rgb_img, dep_img)
        self.align()
        return self.transform.astype(np.float64)

    # ---- cvo.cpp:475-503 ---------------------------------------------------------------------
    def compute_innerproduct(self, tran):
        """-> dict(inn_pre, inn_post, post_hessian, inliers, inn_fixed_pcd, inn_moving_pcd, cos_angle)"""
        a = self.api
        tran = np.asarray(tran, dtype=np.float32)
        if hasattr(a, "compute_innerproduct"):   # one launch for the five queries (CUDA library)
            vals, H, inliers = a.compute_innerproduct(self.h, tran)
            inn_pre, inn_post, inn_fixed, inn_moving = (InnP(*v) for v in vals)
            with np.errstate(divide="ignore", invalid="ignore"):
                cos_angle = np.float32(inn_post.value / (np.sqrt(inn_fixed.value) * np.sqrt(inn_moving.value)))
            return dict(inn_pre=inn_pre, inn_post=inn_post, post_hessian=H, inliers=inliers,
                        inn_fixed_pcd=inn_fixed, inn_moving_pcd=inn_moving, cos_angle=cos_angle)
        inn_pre = InnP(*a.inner_product(self.h, SLOT_MOVING, None, SLOT_FIXED))
        inn_post = InnP(*a.inner_product(self.h, SLOT_MOVING, tran, SLOT_FIXED))
        inn_fixed = InnP(*a.inner_product(self.h, SLOT_FIXED, None, SLOT_FIXED))
        inn_moving = InnP(*a.inner_product(self.h, SLOT_MOVING, None, SLOT_MOVING))
        with np.errstate(divide="ignore", invalid="ignore"):
            cos_angle = np.float32(inn_post.value / (np.sqrt(inn_fixed.value) * np.sqrt(inn_moving.value)))
        H, inliers = a.hessian(self.h, SLOT_MOVING, tran, SLOT_FIXED)
        return dict(inn_pre=inn_pre, inn_post=inn_post, post_hessian=H, inliers=inliers,
                    inn_fixed_pcd=inn_fixed, inn_moving_pcd=inn_moving, cos_angle=cos_angle)

    # ---- cvo.cpp:505-561 ---------------------------------------------------------------------
    def compute_innerproduct_lc(self, prior_tran, lc_prior_tran, lc_prior_tran_2, lc_tran):
        a = self.api
        f32 = lambda m: np.asarray(m, dtype=np.float32)  # noqa: E731
        if hasattr(a, "compute_innerproduct_lc"):   # one launch for the eight queries (CUDA library)
            r = a.compute_innerproduct_lc(self.h, f32(prior_tran), f32(lc_prior_tran), f32(lc_prior_tran_2),
                                          f32(lc_tran))
            inn = [InnP(float(r.value[k]), int(r.num[k])) for k in range(6)]
            return dict(inn_prior=inn[0], inn_lc_prior=inn[1], inn_lc_pre=inn[2], inn_lc_post=inn[3],
                        post_hessian=np.array(r.post_hessian[:], dtype=np.float64).reshape(6, 6),
                        inliers_svd=int(r.inliers_svd), inliers_pnpransac=int(r.inliers_pnpransac),
                        inn_fixed_pcd=inn[4], inn_moving_pcd=inn[5], cos_angle=np.float32(r.cos_angle),
                        accept=bool(r.accept))
        inn_prior = InnP(*a.inner_product(self.h, SLOT_MOVING, f32(prior_tran), SLOT_FIXED))
        inn_lc_prior = InnP(*a.inner_product(self.h, SLOT_MOVING, f32(lc_prior_tran), SLOT_FIXED))
        inn_lc_pre = InnP(*a.inner_product(self.h, SLOT_MOVING, None, SLOT_FIXED))
        inn_lc_post = InnP(*a.inner_product(self.h, SLOT_MOVING, f32(lc_tran), SLOT_FIXED))
        inn_fixed = InnP(*a.inner_product(self.h, SLOT_FIXED, None, SLOT_FIXED))
        inn_moving = InnP(*a.inner_product(self.h, SLOT_MOVING, None, SLOT_MOVING))
        with np.errstate(divide="ignore", invalid="ignore"):
            cos_angle = np.float32(inn_lc_post.value / (np.sqrt(inn_fixed.value) * np.sqrt(inn_moving.value)))
        H, inliers_svd = a.hessian(self.h, SLOT_MOVING, f32(lc_tran), SLOT_FIXED)
        _, inliers_pnpransac = a.hessian(self.h, SLOT_MOVING, f32(lc_prior_tran_2), SLOT_FIXED)
        return dict(inn_prior=inn_prior, inn_lc_prior=inn_lc_prior, inn_lc_pre=inn_lc_pre,
                    inn_lc_post=inn_lc_post, post_hessian=H, inliers_svd=inliers_svd,
                    inliers_pnpransac=inliers_pnpransac, inn_fixed_pcd=inn_fixed,
                    inn_moving_pcd=inn_moving, cos_angle=cos_angle)

    # ---- cvo.cpp:578-618 ---------------------------------------------------------------------
    def update_fixed_pcd(self):
        self.api.slot_move(self.h, SLOT_FIXED, SLOT_MOVING)

    def update_previous_pcd(self):
        self.api.slot_move(self.h, SLOT_PREVIOUS, SLOT_MOVING)
        self.pre_pc_init = True

    def reset_keyframe(self, odometry):
        if not self.pre_pc_init:
            self.api.slot_move(self.h, SLOT_FIXED, SLOT_MOVING)
        else:
            self.api.slot_move(self.h, SLOT_FIXED, SLOT_PREVIOUS)
            self.update_previous_pcd()
        self.reset_transform(odometry)

    def reset_transform(self, odometry):
        self.transform = np.asarray(odometry, dtype=np.float32).copy()

    def reset_initial(self, odometry):
        # the same float operations, in the same order, as include/cvo.hpp (detail::mul44 / inv_affine):
        # the alignment that starts from this prior is sensitive to its last bits
        init = _inv_affine_f32(_mul44_f32(self.transform, np.asarray(odometry, dtype=np.float32)))
        self.api.set_RT(self.h, init[:3, :3], init[:3, 3])
        return _inv_affine_f32(init)

    # ---- getters (cvo.hpp:268-276) -------------------------------------------------------------
    def get_fixed_and_moving_number(self):
        """What set_pcd cached (cvo.cpp:370-371) — also after update_fixed_pcd has moved the clouds, like the
        reference; align() refreshes the cache from its result, so no device sync on the tracking path."""
        if not self._sizes_valid and self.init:
            nf, nm = self.api.slot_size(self.h, SLOT_FIXED), self.api.slot_size(self.h, SLOT_MOVING)
            if nf >= 0 and nm >= 0:
                self.num_fixed, self.num_moving, self._sizes_valid = nf, nm, True
        return self.num_fixed, self.num_moving

    def get_iteration_number(self):
        return self.iter

    def get_A_nonzero(self):
        return self.A_nonzero

    def get_fixed_frame_selected_points(self):
        return self.api.get_selected_points(self.h, SLOT_FIXED)

    def get_moving_frame_selected_points(self):
        return self.api.get_selected_points(self.h, SLOT_MOVING)


def track_sequence(frames, calib, params=None, api=None, device=0, dedup=False):
    """The per-frame call pattern of LocalTracker (src/local_tracker.cpp:223-251, 349-431) with two
    cvo objects (consecutive-frame odometry and keyframe tracking).  Keyframes are never
    replaced here (the keyframe decision lives in KeyframeTracker, out of scope); the first
    frame is the keyframe.  `frames` is a list of (bgr, depth).  Returns per-frame dicts.
    dedup: the keyframe object adopts the cloud the odometry object selected for the same image
    (set_pcd_from / match_keyframe_from) instead of selecting it again (CUDA library only)."""
    odo = Cvo(calib, params, api, device)
    kf = Cvo(calib, params, api, device)
    out = []
    # initNewLocalMap (local_tracker.cpp:223-345): both objects take the keyframe, odometry
    # aligns the second frame, the keyframe object adopts the odometry transform
    odo.set_pcd(*frames[0])
    if dedup:
        kf.set_pcd_from(odo, SLOT_FIXED)
    else:
        kf.set_pcd(*frames[0])
    T = odo.match_odometry(*frames[1])
    r = odo.compute_innerproduct(T.astype(np.float32))
    kf.first_frame = False
    kf.reset_transform(T.astype(np.float32))
    out.append(dict(odometry=T, keyframe=T.copy(), r_odometry=r, r_keyframe=r))
    odo.update_fixed_pcd()
    for rgb, dep in frames[2:]:
        T_odo = odo.match_odometry(rgb, dep)
        r_odo = odo.compute_innerproduct(T_odo.astype(np.float32))
        odo.update_fixed_pcd()
        kf.reset_initial(T_odo.astype(np.float32))
        T_kf = kf.match_keyframe_from(odo, SLOT_FIXED) if dedup else kf.match_keyframe(rgb, dep)
        r_kf = kf.compute_innerproduct(T_kf.astype(np.float32))
        kf.update_previous_pcd()   # accepted frame (local_tracker.cpp:506)
        out.append(dict(odometry=T_odo, keyframe=T_kf, r_odometry=r_odo, r_keyframe=r_kf))
    odo.close()
    kf.close()
    return out
