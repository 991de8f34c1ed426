"""Image ingest (src/run_SLAM.cpp:134-143: cv::imread of the colour and depth PNGs).

CPU: the library's PNG decoder produces the buffers cv::imread produces (checked against cv2 where it is installed,
and against the source arrays for hand-made files that use all five scan-line filters, 8- and 16-bit, RGB / RGBA /
gray); malformed input returns error codes; the prefetch ring hands frames back in order.
GPU: cvo_set_frame_png and the prefetcher feed the selection the same bytes as the raw-buffer path: identical clouds.
"""
import ctypes as C
import struct
import zlib

import numpy as np
import pytest


def _lib():
    from cvo_slam_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    vp, sz, ip = C.c_void_p, C.c_size_t, C.POINTER(C.c_int)
    lib.cvo_png_info.argtypes = [vp, sz, ip, ip, ip, ip]
    lib.cvo_png_decode_bgr8.argtypes = [vp, sz, vp, sz, ip, ip]
    lib.cvo_png_decode_depth16.argtypes = [vp, sz, vp, sz, ip, ip]
    lib.cvo_ingest_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cvo_ingest_destroy.argtypes = [vp]
    lib.cvo_ingest_submit.argtypes = [vp, vp, sz, vp, sz]
    lib.cvo_ingest_wait.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip, ip]
    lib.cvo_ingest_release.argtypes = [vp]
    lib.cvo_set_frame_png.argtypes = [vp, C.c_int, vp, sz, vp, sz]
    lib.cvo_last_error.restype = C.c_char_p
    return lib


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def encode_png(arr, filters=(0, 1, 2, 3, 4), level=6, idat_split=3):
    """arr: [h, w] or [h, w, c] uint8 / uint16 -> PNG bytes; scan-line filters cycle through `filters`; the
    compressed stream is cut into `idat_split` IDAT chunks (real encoders do that)."""
    arr = np.asarray(arr)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    h, w, c = arr.shape
    depth = 16 if arr.dtype == np.uint16 else 8
    color = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = arr.astype(">u2").tobytes() if depth == 16 else arr.astype(np.uint8).tobytes()
    bpp = c * depth // 8
    stride = w * bpp
    rows = np.frombuffer(raw, np.uint8).reshape(h, stride).astype(np.int32)
    out = bytearray()
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = rows[y]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - ((left + prev) >> 1)
        else:
            pa, pb, pc = np.abs(prev - ul), np.abs(left - ul), np.abs(left + prev - 2 * ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            f = cur - pred
        out.append(ft)
        out += (f & 255).astype(np.uint8).tobytes()
        prev = cur
    z = zlib.compress(bytes(out), level)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 0))
    png += chunk(b"tEXt", b"Comment\x00made by tests/test_ingest.py")
    cut = [len(z) * k // idat_split for k in range(idat_split + 1)]
    for k in range(idat_split):
        png += chunk(b"IDAT", z[cut[k]:cut[k + 1]])
    return png + chunk(b"IEND", b"")


def _decode_bgr(lib, png):
    buf = np.frombuffer(png, np.uint8)
    w, h = C.c_int(0), C.c_int(0)
    assert lib.cvo_png_decode_bgr8(buf.ctypes.data, len(png), None, 0, C.byref(w), C.byref(h)) == 0
    out = np.zeros((h.value, w.value, 3), np.uint8)
    rc = lib.cvo_png_decode_bgr8(buf.ctypes.data, len(png), out.ctypes.data, out.nbytes, None, None)
    assert rc == 0, lib.cvo_last_error()
    return out


def _decode_depth(lib, png):
    buf = np.frombuffer(png, np.uint8)
    w, h = C.c_int(0), C.c_int(0)
    assert lib.cvo_png_decode_depth16(buf.ctypes.data, len(png), None, 0, C.byref(w), C.byref(h)) == 0
    out = np.zeros((h.value, w.value), np.uint16)
    rc = lib.cvo_png_decode_depth16(buf.ctypes.data, len(png), out.ctypes.data, out.size, None, None)
    assert rc == 0, lib.cvo_last_error()
    return out


def test_png_decoder_all_filters_and_formats():
    lib = _lib()
    rng = np.random.default_rng(3)
    # smooth + noise content so that every filter type produces non-trivial residuals
    yy, xx = np.mgrid[0:97, 0:131]
    base = (128 + 90 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.int32)
    rgb = np.clip(base[:, :, None] + rng.integers(-20, 20, (97, 131, 3)), 0, 255).astype(np.uint8)
    for filt in ((0,), (1,), (2,), (3,), (4,), (0, 1, 2, 3, 4)):
        assert np.array_equal(_decode_bgr(lib, encode_png(rgb, filt)), rgb[:, :, ::-1])
    rgba = np.concatenate([rgb, rng.integers(0, 255, (97, 131, 1)).astype(np.uint8)], -1)
    assert np.array_equal(_decode_bgr(lib, encode_png(rgba)), rgb[:, :, ::-1])          # alpha dropped
    gray = rgb[:, :, 0]
    assert np.array_equal(_decode_bgr(lib, encode_png(gray)), np.repeat(gray[:, :, None], 3, -1))
    rgb16 = (rgb.astype(np.uint16) << 8) | rng.integers(0, 255, rgb.shape).astype(np.uint16)
    assert np.array_equal(_decode_bgr(lib, encode_png(rgb16)), rgb[:, :, ::-1])         # 16 -> 8 bit keeps the high byte
    depth = (rng.integers(0, 65535, (97, 131))).astype(np.uint16)
    depth[::7, ::5] = 0
    for filt in ((0,), (4,), (0, 1, 2, 3, 4)):
        assert np.array_equal(_decode_depth(lib, encode_png(depth, filt)), depth)
    info = [C.c_int(0) for _ in range(4)]
    png = encode_png(depth)
    buf = np.frombuffer(png, np.uint8)
    assert lib.cvo_png_info(buf.ctypes.data, len(png), *[C.byref(x) for x in info]) == 0
    assert [x.value for x in info] == [131, 97, 1, 16]


def test_png_decoder_matches_cv2_imread_semantics():
    cv2 = pytest.importorskip("cv2")
    lib = _lib()
    rng = np.random.default_rng(4)
    bgr = rng.integers(0, 255, (120, 160, 3)).astype(np.uint8)
    depth = rng.integers(0, 65535, (120, 160)).astype(np.uint16)
    ok, enc = cv2.imencode(".png", bgr)            # libpng's own filter heuristics and chunking
    assert ok
    png = enc.tobytes()
    assert np.array_equal(_decode_bgr(lib, png), cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_COLOR))
    assert np.array_equal(_decode_bgr(lib, png), bgr)
    ok, enc = cv2.imencode(".png", depth)
    png = enc.tobytes()
    assert np.array_equal(_decode_depth(lib, png), cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_ANYDEPTH))
    # our own encoder is read the same way by cv2
    mine = encode_png(bgr[:, :, ::-1])
    assert np.array_equal(cv2.imdecode(np.frombuffer(mine, np.uint8), cv2.IMREAD_COLOR), bgr)


def test_png_decoder_rejects_malformed_input():
    lib = _lib()
    rgb = np.zeros((64, 64, 3), np.uint8)
    png = bytearray(encode_png(rgb))
    out = np.zeros((64, 64, 3), np.uint8)

    def dec(b, cap=out.nbytes):
        buf = np.frombuffer(bytes(b), np.uint8)
        return lib.cvo_png_decode_bgr8(buf.ctypes.data, len(b), out.ctypes.data, cap, None, None)

    assert dec(png) == 0
    assert dec(png[:40]) == -1                      # truncated
    assert dec(b"JFIF" + bytes(png[4:])) == -1      # wrong signature
    bad = bytearray(png)
    bad[-20] ^= 0x55                                # corrupt the last IDAT: CRC mismatch
    assert dec(bad) == -1
    assert dec(png, cap=100) == -4                  # CVO_ERR_CAPACITY
    dpt = encode_png(np.zeros((64, 64), np.uint8))  # 8-bit gray is not a depth map
    buf = np.frombuffer(dpt, np.uint8)
    d16 = np.zeros((64, 64), np.uint16)
    assert lib.cvo_png_decode_depth16(buf.ctypes.data, len(dpt), d16.ctypes.data, d16.size, None, None) == -1
    interlaced = bytearray(png)
    interlaced[28] = 1                              # IHDR interlace byte (CRC now wrong too: rejected at the header)
    assert dec(interlaced) == -1


def test_ingest_ring_returns_frames_in_order():
    lib = _lib()
    rng = np.random.default_rng(5)
    frames = []
    for k in range(7):
        rgb = rng.integers(0, 255, (64, 96, 3)).astype(np.uint8)
        d = rng.integers(0, 65535, (64, 96)).astype(np.uint16)
        frames.append((rgb, d, encode_png(rgb, (k % 5,)), encode_png(d, ((k + 2) % 5,))))
    g = C.c_void_p()
    assert lib.cvo_ingest_create(96, 64, 3, 2, C.byref(g)) == 0
    it = iter(frames)
    pending = 0
    done = 0
    for rgb, d, p1, p2 in frames[:3]:
        b1, b2 = np.frombuffer(p1, np.uint8), np.frombuffer(p2, np.uint8)
        assert lib.cvo_ingest_submit(g, b1.ctypes.data, len(p1), b2.ctypes.data, len(p2)) == 0
        pending += 1
    nxt = 3
    while done < len(frames):
        pb, pd, w, h = C.c_void_p(), C.c_void_p(), C.c_int(0), C.c_int(0)
        assert lib.cvo_ingest_wait(g, C.byref(pb), C.byref(pd), C.byref(w), C.byref(h)) == 0
        got = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint8)), (64, 96, 3)).copy()
        gd = np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint16)), (64, 96)).copy()
        assert (w.value, h.value) == (96, 64)
        assert np.array_equal(got, frames[done][0][:, :, ::-1]) and np.array_equal(gd, frames[done][1])
        assert lib.cvo_ingest_release(g) == 0
        done += 1
        if nxt < len(frames):
            p1, p2 = frames[nxt][2], frames[nxt][3]
            b1, b2 = np.frombuffer(p1, np.uint8), np.frombuffer(p2, np.uint8)
            assert lib.cvo_ingest_submit(g, b1.ctypes.data, len(p1), b2.ctypes.data, len(p2)) == 0
            nxt += 1
    pb, pd = C.c_void_p(), C.c_void_p()
    assert lib.cvo_ingest_wait(g, C.byref(pb), C.byref(pd), None, None) == -3   # nothing submitted
    # a frame of the wrong size is reported, not decoded into the ring
    bad = encode_png(np.zeros((65, 96, 3), np.uint8))
    p2 = frames[0][3]
    b1, b2 = np.frombuffer(bad, np.uint8), np.frombuffer(p2, np.uint8)
    assert lib.cvo_ingest_submit(g, b1.ctypes.data, len(bad), b2.ctypes.data, len(p2)) == 0
    assert lib.cvo_ingest_wait(g, C.byref(pb), C.byref(pd), None, None) != 0
    assert lib.cvo_ingest_release(g) == 0
    lib.cvo_ingest_destroy(g)


@pytest.mark.gpu
def test_set_frame_png_gives_the_clouds_of_the_raw_path(cuda_api, tum_calib, pair_c1):
    lib = _lib()
    bgr_a, d_a, bgr_b, d_b, _ = pair_c1
    h_raw, h_png, h_ring = cuda_api.create(tum_calib), cuda_api.create(tum_calib), cuda_api.create(tum_calib)
    g = C.c_void_p()
    assert lib.cvo_ingest_create(640, 480, 2, 2, C.byref(g)) == 0
    pngs = []
    for bgr, d in ((bgr_a, d_a), (bgr_b, d_b)):
        p1, p2 = encode_png(bgr[:, :, ::-1]), encode_png(d)     # a PNG stores R G B; imread returns B G R
        pngs.append((np.frombuffer(p1, np.uint8), np.frombuffer(p2, np.uint8)))
        assert lib.cvo_ingest_submit(g, pngs[-1][0].ctypes.data, len(p1), pngs[-1][1].ctypes.data, len(p2)) == 0
    for slot, (bgr, d) in enumerate(((bgr_a, d_a), (bgr_b, d_b))):
        cuda_api.set_frame(h_raw, slot, bgr, d)
        b1, b2 = pngs[slot]
        assert lib.cvo_set_frame_png(h_png, slot, b1.ctypes.data, b1.size, b2.ctypes.data, b2.size) == 0, lib.cvo_last_error()
        pb, pd, w, hh = C.c_void_p(), C.c_void_p(), C.c_int(0), C.c_int(0)
        assert lib.cvo_ingest_wait(g, C.byref(pb), C.byref(pd), C.byref(w), C.byref(hh)) == 0
        assert cuda_api.lib.cvo_set_frame(h_ring, slot, pb, 640 * 3, pd, 640 * 2, 640, 480) == 0
        cuda_api.slot_size(h_ring, slot)     # (the copy from the pinned slot has completed before the slot is released)
        assert lib.cvo_ingest_release(g) == 0
    for slot in (0, 1):
        p0, f0 = cuda_api.get_cloud(h_raw, slot)
        for hx in (h_png, h_ring):
            p1, f1 = cuda_api.get_cloud(hx, slot)
            assert np.array_equal(p0.view(np.uint32), p1.view(np.uint32)) and np.array_equal(f0.view(np.uint32), f1.view(np.uint32))
            assert np.array_equal(cuda_api.get_selected_points(h_raw, slot), cuda_api.get_selected_points(hx, slot))
    r0, _ = cuda_api.align(h_raw)
    r1, _ = cuda_api.align(h_png)
    assert np.array_equal(r0.transform_np(), r1.transform_np())
    lib.cvo_ingest_destroy(g)
    for hx in (h_raw, h_png, h_ring):
        cuda_api.destroy(hx)
