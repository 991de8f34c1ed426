"""Batches of independent frame pairs (the loop-closure verification pattern of
src/keyframe_graph.cpp:622-731: a fresh cvo object per candidate, `reset_initial` prior,
set_pcd x2, align, inner product) through the C ABI's cvo_batch_* entry points.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class Batch:
    def __init__(self, calib, params=None, max_frames=64, max_pairs=64, width=640, height=480,
                 device=0, api=None):
        self.api = api if api is not None else capi.load()
        self.lib = self.api.lib
        self.params = params if params is not None else self.api.default_params()
        self.w, self.h = width, height
        self.max_frames, self.max_pairs = max_frames, max_pairs
        self.b = C.c_void_p()
        self.api._check(self.lib.cvo_batch_create(C.byref(calib), C.byref(self.params), device,
                                                  max_frames, max_pairs, width, height,
                                                  C.byref(self.b)), "batch_create")

    def close(self):
        if self.b:
            self.lib.cvo_batch_destroy(self.b)
            self.b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- frames ------------------------------------------------------------------------------
    def set_frames(self, bgr, depth, first=0):
        """bgr [n,h,w,3] uint8, depth [n,h,w] uint16 host arrays (pinned memory makes the copies async)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        n = bgr.shape[0]
        assert bgr.shape == (n, self.h, self.w, 3) and depth.shape == (n, self.h, self.w)
        # the call only enqueues the copies: the arrays must outlive it (until a consuming call has returned)
        self._keep = getattr(self, "_keep", [])[-3:] + [(bgr, depth)]
        self.api._check(self.lib.cvo_batch_set_frames(self.b, first, n, bgr.ctypes.data,
                                                      depth.ctypes.data), "batch_set_frames")

    def set_frames_ptr(self, bgr_ptr, depth_ptr, n, first=0, device=False):
        fn = self.lib.cvo_batch_set_frames_device if device else self.lib.cvo_batch_set_frames
        self.api._check(fn(self.b, first, n, bgr_ptr, depth_ptr), "batch_set_frames")

    def frame_size(self, k):
        n = C.c_int(0)
        self.api._check(self.lib.cvo_batch_frame_size(self.b, k, C.byref(n)), "batch_frame_size")
        return n.value

    # ---- pairs -------------------------------------------------------------------------------
    def make_pairs(self, pairs, R=None, T=None, ell=None):
        """pairs: iterable of (fixed_frame, moving_frame); optional per-pair initial R [n,3,3],
        T [n,3] (e.g. from reset_initial) and ell.  Fresh-object defaults otherwise."""
        pairs = np.asarray(list(pairs), dtype=np.int32).reshape(-1, 2)
        n = len(pairs)
        d = np.zeros(n, dtype=capi.PAIR_DTYPE)
        d["fixed_frame"] = pairs[:, 0]
        d["moving_frame"] = pairs[:, 1]
        d["R"] = np.eye(3, dtype=np.float32).reshape(9) if R is None else np.asarray(R, np.float32).reshape(n, 9)
        d["T"] = 0 if T is None else np.asarray(T, np.float32).reshape(n, 3)
        d["ell"] = self.params.ell_init if ell is None else ell
        return d

    def align(self, pairs, R=None, T=None, ell=None):
        d = pairs if isinstance(pairs, np.ndarray) and pairs.dtype == capi.PAIR_DTYPE \
            else self.make_pairs(pairs, R, T, ell)
        res = np.zeros(len(d), dtype=capi.RESULT_DTYPE)
        self.api._check(self.lib.cvo_batch_align(self.b, len(d), d.ctypes.data, res.ctypes.data),
                        "batch_align")
        self._last_pairs = d
        return res

    def inner_product(self, pairs, results):
        d = pairs if isinstance(pairs, np.ndarray) and pairs.dtype == capi.PAIR_DTYPE \
            else self.make_pairs(pairs)
        vals = np.zeros(len(d), np.float32)
        nums = np.zeros(len(d), np.int32)
        res = np.ascontiguousarray(results)
        self.api._check(self.lib.cvo_batch_inner_product(self.b, len(d), d.ctypes.data, res.ctypes.data,
                                                         vals.ctypes.data, nums.ctypes.data),
                        "batch_inner_product")
        return vals, nums

    def verify_lc(self, pairs, results, prior_tran, lc_prior_tran, lc_prior_tran_2):
        """compute_innerproduct_lc (cvo.cpp:505-561) + the accept rule of keyframe_graph.cpp:711-712 for
        every pair in one launch; the three priors are [n,4,4] float32, lc_tran is results.transform.
        -> structured array of capi.LC_DTYPE"""
        d = pairs if isinstance(pairs, np.ndarray) and pairs.dtype == capi.PAIR_DTYPE \
            else self.make_pairs(pairs)
        n = len(d)
        res = np.ascontiguousarray(results)
        pr = [np.ascontiguousarray(np.asarray(m, np.float32).reshape(n, 16)) for m in
              (prior_tran, lc_prior_tran, lc_prior_tran_2)]
        out = np.zeros(n, dtype=capi.LC_DTYPE)
        self.api._check(self.lib.cvo_batch_verify_lc(self.b, n, d.ctypes.data, res.ctypes.data, pr[0].ctypes.data,
                                                     pr[1].ctypes.data, pr[2].ctypes.data, out.ctypes.data),
                        "batch_verify_lc")
        return out

    def stats(self):
        s = (C.c_int64 * 4)()
        self.api._check(self.lib.cvo_batch_stats(self.b, s), "batch_stats")
        return dict(launches=s[0], evals=s[1], iterations=s[2], nnz=s[3])

    def phase_cycles(self):
        """cumulative SM cycles (thread 0 of every CTA) per phase of the align kernel, and list rebuilds"""
        s = (C.c_int64 * 8)()
        self.lib.cvo_batch_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        self.api._check(self.lib.cvo_batch_phase_cycles(self.b, s), "batch_phase_cycles")
        names = ["grid", "P0", "P1a_search", "P1b", "P2", "P3", "filters", "rebuilds"]
        return {k: int(s[i]) for i, k in enumerate(names)}

    def mark(self, which):
        self.api._check(self.lib.cvo_batch_mark(self.b, which), "batch_mark")

    def elapsed_ms(self):
        ms = C.c_float(0)
        self.api._check(self.lib.cvo_batch_elapsed_ms(self.b, C.byref(ms)), "batch_elapsed_ms")
        return ms.value

    def last_align_ms(self):
        ms = C.c_float(0)
        self.api._check(self.lib.cvo_batch_last_align_ms(self.b, C.byref(ms)), "batch_last_align_ms")
        return ms.value


class MultiBatch:
    """cvo_multi_*: one list of host frames and pairs over several GPUs of this process (contiguous
    blocks of pairs per device, one host thread per device, no exchange between devices)."""

    def __init__(self, calib, params=None, n_devices=1, devices=None, max_frames=64, max_pairs=64,
                 width=640, height=480, api=None):
        self.api = api if api is not None else capi.load()
        self.lib = self.api.lib
        self.params = params if params is not None else self.api.default_params()
        self.n_devices, self.w, self.h = n_devices, width, height
        dv = None if devices is None else (C.c_int * n_devices)(*devices)
        self.m = C.c_void_p()
        self.api._check(self.lib.cvo_multi_create(C.byref(calib), C.byref(self.params), n_devices, dv,
                                                  max_frames, max_pairs, width, height, C.byref(self.m)),
                        "multi_create")

    def close(self):
        if self.m:
            self.lib.cvo_multi_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align(self, bgr, depth, pairs, inner_products=True):
        """bgr [n,h,w,3] uint8, depth [n,h,w] uint16 (host), pairs: PAIR_DTYPE array -> (results, values, nums)"""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        d = np.ascontiguousarray(pairs)
        assert d.dtype == capi.PAIR_DTYPE
        res = np.zeros(len(d), dtype=capi.RESULT_DTYPE)
        vals = np.zeros(len(d), np.float32) if inner_products else None
        nums = np.zeros(len(d), np.int32) if inner_products else None
        self.api._check(self.lib.cvo_multi_align(self.m, bgr.shape[0], bgr.ctypes.data, depth.ctypes.data, len(d),
                                                 d.ctypes.data, res.ctypes.data,
                                                 vals.ctypes.data if inner_products else None,
                                                 nums.ctypes.data if inner_products else None), "multi_align")
        return res, vals, nums

    def last_shares(self):
        f = (C.c_int * self.n_devices)()
        p = (C.c_int * self.n_devices)()
        ms = (C.c_float * self.n_devices)()
        self.api._check(self.lib.cvo_multi_last_shares(self.m, f, p, ms), "multi_last_shares")
        return dict(frames=list(f), pairs=list(p), align_ms=[float(x) for x in ms])
