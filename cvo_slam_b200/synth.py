"""Deterministic procedural RGB-D frames of TUM / ETH3D shape (SURVEY §8d "Synthetic inputs").

An analytic room (axis-aligned walls, floor, ceiling) with a few boxes standing on the
floor is ray-cast from a pinhole camera.  Colour is a *solid* (3-D) texture — a sum of
sinusoids per channel plus constant-colour cuboids with sharp edges — so that two views
of the scene are photometrically and geometrically consistent under a known SE(3).
Sensor noise N(0, 2) per channel and 5 % depth holes are added per frame.

Everything is torch, so the same code renders on the CPU (tests, golden fixtures) and on
the GPU (bench: thousands of frames).  Poses are camera-to-world 4x4 matrices; the CVO
`transform` for (fixed=a, moving=b) is inv(T_wa) @ T_wb (moving -> fixed coordinates).
"""
from __future__ import annotations

import math

import numpy as np
import torch

ROOM_LO = (-2.6, -1.6, -1.5)   # x right, y down, z forward (camera convention)
ROOM_HI = (2.6, 1.25, 2.5)     # back wall at z = 2.5 m, floor at y = 1.25 m


def make_scene(seed, n_sin=12, n_cuboids=100, high_gradient=False):
    g = torch.Generator().manual_seed(int(seed))
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)  # noqa: E731
    n_box = 5
    boxes_lo, boxes_hi = [], []
    for i in range(n_box):
        cx = -1.6 + 0.8 * i + 0.3 * (r(1).item() - 0.5)
        cz = 1.3 + 0.9 * r(1).item()
        sx, sz = 0.15 + 0.2 * r(1).item(), 0.15 + 0.2 * r(1).item()
        hy = 0.8 + 0.9 * r(1).item()
        boxes_lo.append([cx - sx, ROOM_HI[1] - hy, cz - sz])
        boxes_hi.append([cx + sx, ROOM_HI[1], cz + sz])
    fmax = 9.0 if high_gradient else 4.0
    # spatial frequencies in cycles / metre, log-uniform magnitude, random direction
    mag = torch.exp(math.log(0.6) + r(3, n_sin) * (math.log(fmax) - math.log(0.6)))
    dirs = torch.randn(3, n_sin, 3, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    amp = (22.0 if high_gradient else 16.0) * (0.4 + r(3, n_sin))
    lo = torch.tensor(ROOM_LO, dtype=torch.float64) - 0.2
    hi = torch.tensor(ROOM_HI, dtype=torch.float64) + 0.2
    ccen = lo + (hi - lo) * r(n_cuboids, 3)
    csize = 0.15 + (0.3 if high_gradient else 0.45) * r(n_cuboids, 3)
    scene = dict(
        boxes_lo=torch.tensor(boxes_lo, dtype=torch.float64),
        boxes_hi=torch.tensor(boxes_hi, dtype=torch.float64),
        freq=(mag.unsqueeze(-1) * dirs) * (2.0 * math.pi),
        phase=2.0 * math.pi * r(3, n_sin),
        amp=amp,
        base=90.0 + 70.0 * r(3),
        cub_lo=ccen - csize,
        cub_hi=ccen + csize,
        cub_col=20.0 + 215.0 * r(n_cuboids, 3),
    )
    return scene


def pose(rotvec=(0.0, 0.0, 0.0), trans=(0.0, 0.0, 0.0)):
    """camera-to-world 4x4 from an axis-angle vector (rad) and a translation (m)."""
    w = np.asarray(rotvec, dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        R = np.eye(3)
    else:
        R = np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = trans
    return T


def _slab(o, d, lo, hi):
    """ray / AABB: (t_enter, t_exit) for rays o + t d, broadcasting over leading dims."""
    inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    t0 = (lo - o) * inv
    t1 = (hi - o) * inv
    tmin = torch.minimum(t0, t1).amax(dim=-1)
    tmax = torch.maximum(t0, t1).amin(dim=-1)
    return tmin, tmax


def render(scene, T_wc, calib, w, h, noise_seed=0, device="cpu", noise_sigma=2.0, hole_frac=0.05):
    """-> (bgr uint8 [h,w,3], depth uint16 [h,w]) on `device`."""
    dev = torch.device(device)
    f64 = torch.float64
    S = {k: v.to(dev) for k, v in scene.items()}
    Twc = torch.as_tensor(np.asarray(T_wc), dtype=f64, device=dev)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev, dtype=f64),
                            torch.arange(w, device=dev, dtype=f64), indexing="ij")
    dc = torch.stack([(xs - calib.cx) / calib.fx, (ys - calib.cy) / calib.fy, torch.ones_like(xs)], -1)
    d = dc @ Twc[:3, :3].T            # world directions (z_cam = 1 => t is depth)
    o = Twc[:3, 3]
    lo = torch.tensor(ROOM_LO, dtype=f64, device=dev)
    hi = torch.tensor(ROOM_HI, dtype=f64, device=dev)
    _, t_room = _slab(o, d, lo, hi)
    t = t_room
    for b in range(S["boxes_lo"].shape[0]):
        te, tx = _slab(o, d, S["boxes_lo"][b], S["boxes_hi"][b])
        hit = (te < tx) & (te > 1e-3)
        t = torch.where(hit & (te < t), te, t)
    p = o + t.unsqueeze(-1) * d       # world hit points [h,w,3]
    # solid texture
    arg = torch.einsum("hwk,csk->hwcs", p, S["freq"]) + S["phase"]
    col = S["base"] + (S["amp"] * torch.sin(arg)).sum(-1)
    # slow shading term so that large planes are not flat
    col = col * (0.85 + 0.15 * torch.sin(0.9 * p[..., 0:1] + 1.3 * p[..., 1:2] + 0.7 * p[..., 2:3]))
    for c in range(S["cub_lo"].shape[0]):
        inside = ((p > S["cub_lo"][c]) & (p < S["cub_hi"][c])).all(-1)
        col = torch.where(inside.unsqueeze(-1), S["cub_col"][c].expand_as(col), col)
    g = torch.Generator(device=dev).manual_seed(int(noise_seed) * 7919 + 17)
    col = col + noise_sigma * torch.randn(col.shape, generator=g, device=dev, dtype=f64)
    bgr = col.round().clamp(0, 255).to(torch.uint8)
    z = (t * calib.scaling_factor).round().clamp(0, 65535)
    holes = torch.rand(z.shape, generator=g, device=dev, dtype=f64) < hole_frac
    z = torch.where(holes, torch.zeros_like(z), z)
    depth = z.to(torch.int32).to(torch.uint16)
    return bgr, depth


def to_numpy(bgr, depth):
    return np.ascontiguousarray(bgr.cpu().numpy()), np.ascontiguousarray(depth.cpu().numpy())


def relative_transform(T_wa, T_wb):
    """ground-truth CVO `transform` for (fixed=a, moving=b): moving -> fixed coordinates."""
    return np.linalg.inv(np.asarray(T_wa)) @ np.asarray(T_wb)


def pose_error(T_est, T_ref):
    """(rotation angle in rad, translation distance in m) between two 4x4 transforms."""
    E = np.linalg.inv(np.asarray(T_ref, np.float64)) @ np.asarray(T_est, np.float64)
    ang = float(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))
    return ang, float(np.linalg.norm(E[:3, 3]))


def make_pair(seed, calib, w=640, h=480, rot_deg=1.0, axis=(0.3, 1.0, 0.2), trans=(0.02, -0.01, 0.015),
              high_gradient=False, device="cpu", noise_sigma=2.0):
    """C1-style pair: frame a at the origin pose, frame b offset by a known SE(3).

    Returns (bgr_a, depth_a, bgr_b, depth_b, T_gt) as numpy; T_gt = moving(b) -> fixed(a)."""
    scene = make_scene(seed, high_gradient=high_gradient)
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / np.linalg.norm(ax)
    T_wa = pose()
    T_wb = pose(ax * math.radians(rot_deg), trans)
    a = to_numpy(*render(scene, T_wa, calib, w, h, noise_seed=seed * 2 + 1, device=device,
                         noise_sigma=noise_sigma))
    b = to_numpy(*render(scene, T_wb, calib, w, h, noise_seed=seed * 2 + 2, device=device,
                         noise_sigma=noise_sigma))
    return a[0], a[1], b[0], b[1], relative_transform(T_wa, T_wb)


def trajectory(n, seed, max_step_m=0.02, max_step_deg=1.0):
    """Smooth camera path (C2): n camera-to-world poses, per-frame motion below the bounds."""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    ph = rng.uniform(0, 2 * np.pi, 6)
    per = rng.uniform(60, 140, 6)
    # amplitudes chosen so that the derivative stays below the per-frame bound
    at = 0.6 * max_step_m * per[:3] / (2 * np.pi)
    ar = 0.6 * math.radians(max_step_deg) * per[3:] / (2 * np.pi)
    poses = []
    for i in k:
        t = at * np.sin(2 * np.pi * i / per[:3] + ph[:3]) * np.array([1.0, 0.5, 0.7])
        r = ar * np.sin(2 * np.pi * i / per[3:] + ph[3:]) * np.array([0.6, 1.0, 0.4])
        poses.append(pose(r, t))
    return poses
