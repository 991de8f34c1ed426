#!/usr/bin/env python
"""bench.py — CVO frame-pair alignments/s on synthetic 640x480 RGB-D keyframe pairs.

Workload (BASELINE.json configs[4], the configuration the multi-GPU metric is quoted on; it is
the largest single-GPU configuration and the one that shards): F keyframes of one synthetic
scene (TUM fr1 intrinsics), P = 8 F independent keyframe pairs with a perturbed prior
(`reset_initial`), as in loop-closure verification (src/keyframe_graph.cpp:622-731).
One *step* = point selection + features for the frames, alignment of the pairs, and the
post-alignment inner product <T*moving, fixed> of every pair.  The job is ONE fixed list of P pairs
over F frames (BASELINE configs[4]: "8192 independent keyframe-pair alignments sharded across
1/2/4/8 B200"): with N ranks every rank takes a contiguous block of P/N pairs, selects only the
frames that block touches, aligns, and the results are gathered on rank 0 — strong scaling, no
collective on the data path.  `value` = P / max-over-ranks step time.  `--scaling weak` keeps the
round-1 variant (every rank its own F frames / P pairs) and is reported as a side object at N > 1.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA path (libcvo_b200.so)
  python bench.py --impl reference ...                           # CPU reference arm (oracle)

`value`: inputs resident in HBM when the timed region starts, timed with CUDA events on the
stream the kernels run on.  `e2e`: the same step through the C ABI with HOST (pinned) images,
H2D of every frame and D2H of every result inside the timed region; every step uploads its own
frames, and the upload of step k+1 is enqueued (cvo_batch_set_frames only enqueues) into the other
half of the arena before step k is aligned, so it runs beside that alignment.  All K uploads, the
first one included, are inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 640, 480
PARTNER_OFFSETS = [1, 2, 3, 5, 8, 13, 21, 34]
FLOP_PER_EVAL = 28          # d2 8 + d2c 14 + 2 scales + products/compare 4
FLOP_PER_NNZ_ITER = 90      # flow 24 + step-size ~66 (cvo.cpp:213-223, 282-306)
SM_COUNT, FP32_LANES, MUFU_LANES = 148, 128, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cvo_b200", choices=["cvo_b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="keyframes per GPU")
    ap.add_argument("--partners", type=int, default=8, help="pairs per keyframe")
    ap.add_argument("--cpu-pairs", type=int, default=24, help="pairs in the bounded CPU sample")
    ap.add_argument("--exp-mode", type=int, default=0, help="0 exact (bit-faithful), 1 MUFU fast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: one list of frames*partners pairs sharded over the ranks (default); "
                         "weak: every rank its own list")
    ap.add_argument("--side-legs", type=int, default=1, help="C1 / C3 / C4 single-pair legs and the fast-mode roofline (N=1 only)")
    ap.add_argument("--sequence-frames", type=int, default=300,
                    help="C2: frames of the sequential-tracking side measurement (N=1 only; 0 = skip)")
    return ap.parse_args()


# ---- workload ------------------------------------------------------------------------------------
def keyframe_poses(n, seed):
    from cvo_slam_b200 import synth
    rng = np.random.default_rng(seed)
    return [synth.pose(rng.uniform(-0.035, 0.035, 3), rng.uniform(-0.05, 0.05, 3)) for _ in range(n)]


def pair_list(n_frames, partners):
    offs = PARTNER_OFFSETS[:partners]
    return [(i, (i + o) % n_frames) for i in range(n_frames) for o in offs]


def pair_priors(pairs, poses, seed):
    """initial (R, T) per pair: the inverse of a perturbed ground-truth transform, i.e. what
    reset_initial (cvo.cpp:611-618) leaves in R, T from a PnP prior."""
    from cvo_slam_b200 import synth
    rng = np.random.default_rng(seed)
    R = np.zeros((len(pairs), 9), np.float32)
    T = np.zeros((len(pairs), 3), np.float32)
    gts = []
    for k, (fi, mi) in enumerate(pairs):
        gt = synth.relative_transform(poses[fi], poses[mi])
        prior = gt @ synth.pose(rng.normal(0, 0.004, 3), rng.normal(0, 0.006, 3))
        M = np.linalg.inv(prior)
        R[k] = M[:3, :3].astype(np.float32).reshape(9)
        T[k] = M[:3, 3].astype(np.float32)
        gts.append(gt)
    return R, T, gts


def render_frames(frame_ids, poses, seed, device, noise_salt=0):
    """-> (bgr uint8 [n,H,W,3], depth uint16 [n,H,W]) torch tensors on `device`."""
    import torch
    from cvo_slam_b200 import capi, synth
    cal = capi.TUM1_CALIB()
    scene = synth.make_scene(seed)
    bgr = torch.empty((len(frame_ids), H, W, 3), dtype=torch.uint8, device=device)
    dep = torch.empty((len(frame_ids), H, W), dtype=torch.int16, device=device)   # uint16 bit pattern
    for k, f in enumerate(frame_ids):
        b, d = synth.render(scene, poses[f], cal, W, H, noise_seed=seed * 100003 + f + 7919 * noise_salt, device=device)
        bgr[k] = b
        dep[k] = d.view(torch.int16)
    return bgr, dep


# ---- clocks --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---- CPU reference (oracle) leg --------------------------------------------------------------------
def cpu_leg(pairs, poses, R0, T0, seed, n_pairs, steps, warmup, device, partners=8):
    """Times the CPU restatement of the reference on a bounded sample of the same workload, with all
    host threads (torchrun exports OMP_NUM_THREADS=1: overridden here).  The workload selects every
    keyframe ONCE and aligns `partners` pairs per keyframe, so the sample times the two stages
    separately — selection of the sample's keyframes, alignment + inner product of the sample's pairs —
    and combines them at the workload's ratio: seconds per pair = t_align + t_select / partners.
    Returns alignments/s and a description."""
    from cvo_slam_b200 import capi
    from oracle import oracle
    orc = oracle.load()
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.set_num_threads(ncpu)
    cores = orc.num_threads()
    assert cores > 1 or ncpu == 1, f"CPU arm is running on {cores} thread(s) of {ncpu}"
    cal = capi.TUM1_CALIB()
    sample = pairs[:n_pairs]
    frame_ids = sorted({f for p in sample for f in p})
    bgr, dep = render_frames(frame_ids, poses, seed, device)
    bgr, dep = bgr.cpu().numpy(), dep.cpu().numpy().view(np.uint16)
    local = {f: k for k, f in enumerate(frame_ids)}
    t_sel, t_ali = [], []
    for it in range(warmup + steps):
        clouds = {}
        h = orc.create(cal)
        t0 = time.perf_counter()
        for f in frame_ids:
            orc.set_frame(h, 0, bgr[local[f]], dep[local[f]])
        t1 = time.perf_counter()
        for f in frame_ids:   # (untimed) keep the clouds: pairs reuse them, as in the workload
            orc.set_frame(h, 0, bgr[local[f]], dep[local[f]])
            clouds[f] = orc.get_cloud(h, 0)
        t2 = time.perf_counter()
        for k, (fi, mi) in enumerate(sample):
            orc.set_cloud(h, 0, *clouds[fi])
            orc.set_cloud(h, 1, *clouds[mi])
            orc.set_ell(h, 0.15)
            orc.set_RT(h, R0[k].reshape(3, 3), T0[k])
            res, _ = orc.align(h)
            orc.inner_product(h, 1, res.transform_np(), 0)
        t3 = time.perf_counter()
        orc.destroy(h)
        if it >= warmup:
            t_sel.append((t1 - t0) / len(frame_ids))
            t_ali.append((t3 - t2) / len(sample))
    sel, ali = sum(t_sel) / len(t_sel), sum(t_ali) / len(t_ali)
    per_pair = ali + sel / partners
    return dict(value=1.0 / per_pair, unit="alignments/s", cores=cores,
                kind="port",
                select_ms_per_frame=sel * 1e3, align_ms_per_pair=ali * 1e3,
                sample=f"{len(sample)} pairs and {len(frame_ids)} keyframes of the same workload, {steps} timed "
                       f"repetition(s); selection and alignment timed separately and combined at the workload's "
                       f"ratio of 1 selection per {partners} pairs; oracle/"
                       f"{'_ref nanoflann KD-tree' if orc.has_nanoflann else 'cell-list'} radius search, "
                       f"OpenMP over points, {cores} threads"), per_pair * len(sample) * 1e3


def committed_ncu_traffic(n_frames, partners, exp_mode):
    """DRAM bytes (read + write) of one k_align_batch launch from the committed `ncu --set full` summary of
    THIS workload and mode (profiles/r02z_align_batch_{exact,fast}_full.txt: bench.py defaults, 8 192 pairs),
    or None for any other configuration — a number taken under the profiler is never scaled to another size."""
    if (n_frames, partners) != (1024, 8) or exp_mode not in (0, 1):
        return None, None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                        "r02z_align_batch_%s_full.txt" % ("exact" if exp_mode == 0 else "fast"))
    try:
        txt = open(path).read()
    except OSError:
        return None, None
    import re
    rd = re.search(r"dram__bytes_read\.sum \[Gbyte\] = ([0-9.]+)", txt)
    wr = re.search(r"dram__bytes_write\.sum \[Gbyte\] = ([0-9.]+)", txt)
    if not (rd and wr):
        return None, None
    return (float(rd.group(1)) + float(wr.group(1))) * 1e9, os.path.relpath(path, os.path.dirname(os.path.abspath(__file__)))


def measured_hbm_peak():
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6549.0, "fallback: 6549 GB/s (the pool's measured copy bandwidth at the time of writing)"


def cpp_sequence(frames, T_kf_last_python, dedup=False):
    """Compiles scripts/seq_dropin.cpp against include/cvo.hpp + libcvo_b200.so and runs it on `frames`
    (list of (bgr, depth) arrays).  -> dict(ms_per_frame, passes, same_bits) or None."""
    from cvo_slam_b200 import capi
    root = os.path.dirname(os.path.abspath(__file__))
    libdir = os.path.dirname(capi.LIB_PATH)
    try:
        with tempfile.TemporaryDirectory() as td:
            exe = os.path.join(td, "seq_dropin")
            subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "include"),
                            os.path.join(root, "scripts", "seq_dropin.cpp"), "-o", exe, "-L", libdir, "-lcvo_b200",
                            f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
            np.stack([f[0] for f in frames]).tofile(os.path.join(td, "bgr.raw"))
            np.stack([f[1] for f in frames]).astype(np.uint16).tofile(os.path.join(td, "depth.raw"))
            calib = os.path.join(td, "calib.yaml")
            with open(calib, "w") as f:   # config/TUM1.yaml:8-20
                f.write("%YAML:1.0\nCamera.fx: 517.306408\nCamera.fy: 516.469215\nCamera.cx: 318.643040\n"
                        "Camera.cy: 255.313989\nDepthMapFactor: 5000.0\n")
            out = subprocess.run([exe, calib, os.path.join(td, "bgr.raw"), os.path.join(td, "depth.raw"),
                                  str(len(frames)), str(W), str(H), "3", "1" if dedup else "0"], check=True, capture_output=True, text=True).stdout
        tok = out.split()
        i = tok.index("ms_per_frame")
        j = tok.index("T_kf_last")
        passes = [float(x) for x in tok[i + 1:j]]
        T = np.array([float(x) for x in tok[j + 1:j + 17]], np.float32).reshape(4, 4)
        return dict(ms_per_frame=min(passes), passes=[round(x, 3) for x in passes],
                    same_bits=bool(np.array_equal(T, np.asarray(T_kf_last_python, np.float32))))
    except Exception as e:   # no compiler on the box, or the runner failed: the Python leg stands
        print("cpp_sequence unavailable:", repr(e)[:300], file=sys.stderr)
        return None


def sequence_leg(n_frames, api, device, cpu_frames=6):
    """BASELINE configs[1]: a TUM-shaped sequence tracked frame by frame with the LocalTracker call
    pattern (two cvo objects, persistent R/T/ell; 2 set_pcd + 2 align + 2 compute_innerproduct per
    frame).  A dependency chain: one GPU, latency-bound.  Returns frames/s and alignments/s from host
    images (H2D inside), beside the CPU oracle on the first few frames."""
    from cvo_slam_b200 import capi, cvo as cvo_mod, synth
    cal = capi.TUM1_CALIB()
    scene = synth.make_scene(2)
    poses = synth.trajectory(n_frames, 2)
    frames = []
    for k, P in enumerate(poses):
        b, d = synth.render(scene, P, cal, W, H, noise_seed=20 + k, device=device)
        frames.append(synth.to_numpy(b, d))
    cvo_mod.track_sequence(frames[:4], cal, api=api)            # warm-up (allocations, attributes)
    # a latency chain timed from the host is sensitive to whatever else the host does: three passes
    # over the sequence, the fastest is reported (all three are listed)
    passes = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = cvo_mod.track_sequence(frames, cal, api=api)
        passes.append(time.perf_counter() - t0)
    dt = min(passes)
    err = [synth.pose_error(o["keyframe"], synth.relative_transform(poses[0], poses[k + 1])) for k, o in enumerate(out)]
    res = dict(workload=f"C2: {n_frames}-frame synthetic TUM-shaped sequence, LocalTracker call pattern, 1 GPU",
               frames_per_s=(n_frames - 1) / dt, alignments_per_s=(2 * (n_frames - 1) - 1) / dt,
               ms_per_frame=dt / (n_frames - 1) * 1e3,
               ms_per_frame_passes=[round(x / (n_frames - 1) * 1e3, 3) for x in passes],
               max_keyframe_pose_error=dict(rad=float(max(e[0] for e in err)), m=float(max(e[1] for e in err))))
    # the same sequence through the product's own host side: the drop-in C++ class (include/cvo.hpp)
    # driven by scripts/seq_dropin.cpp exactly like LocalTracker drives the reference's class.  No
    # Python between the calls: this is the number a CVO-SLAM build linked against libcvo_b200.so sees.
    cpp = cpp_sequence(frames, out[-1]["keyframe"])
    if cpp:
        res["python_harness_ms_per_frame"] = res["ms_per_frame"]
        res["python_harness_ms_per_frame_passes"] = res.pop("ms_per_frame_passes")
        res.update(host="C++ drop-in class include/cvo.hpp (scripts/seq_dropin.cpp), fastest of three passes",
                   ms_per_frame=cpp["ms_per_frame"], ms_per_frame_passes=cpp["passes"],
                   frames_per_s=1e3 / cpp["ms_per_frame"],
                   alignments_per_s=(2 * (n_frames - 1) - 1) / ((n_frames - 1) * cpp["ms_per_frame"] * 1e-3),
                   cpp_matches_python_bits=cpp["same_bits"])
        dd = cpp_sequence(frames, out[-1]["keyframe"], dedup=True)
        if dd:   # opt-in de-duplicated front end: one selection per frame instead of two
            res["dedup_front_end"] = dict(ms_per_frame=dd["ms_per_frame"], frames_per_s=1e3 / dd["ms_per_frame"],
                                          same_bits_as_two_selections=dd["same_bits"])
    else:
        res["host"] = "Python ctypes mirror of the class (cvo_slam_b200/cvo.py); the C++ runner could not be built here"
    if cpu_frames:
        from oracle import oracle
        orc = oracle.load()
        t0 = time.perf_counter()
        cvo_mod.track_sequence(frames[:cpu_frames], cal, api=orc)
        dtc = time.perf_counter() - t0
        res["cpu_port_frames_per_s"] = (cpu_frames - 1) / dtc
        res["cpu_cores"] = orc.num_threads()
    return res


def single_pair_legs(api, sm_mhz):
    """Side legs at N = 1 (BASELINE configs[0], [2], [3]): one pair on one handle — C1 (640x480, 1 deg / 2.7 cm),
    C3 (dense selection, ~18 k points) and C4 (739x458, 8 deg / 0.15 m, default and ell_init 0.25) — exact mode:
    host-timed cvo_align (launch + D2H of the result), kernel evaluations per second and the FP32 fraction."""
    from cvo_slam_b200 import capi, synth
    fp32_peak = SM_COUNT * FP32_LANES * 2 * sm_mhz * 1e6
    eval_roof = min(fp32_peak / FLOP_PER_EVAL, SM_COUNT * MUFU_LANES * sm_mhz * 1e6 / 2)
    out = {}

    def leg(name, cal, a, da, b, db, params, reps, note):
        h = api.create(cal, params)
        api.set_frame(h, 0, a, da)
        api.set_frame(h, 1, b, db)
        n = (api.slot_size(h, 0), api.slot_size(h, 1))
        best = None
        for rep in range(reps + 1):   # first repetition = warm-up
            api.set_RT(h, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
            api.set_ell(h, params.ell_init)
            s0 = api.handle_stats(h)
            t0 = time.perf_counter()
            res, _recs = api.align(h)
            dt = time.perf_counter() - t0
            s1 = api.handle_stats(h)
            if rep and (best is None or dt < best[0]):
                best = (dt, s1["evals"] - s0["evals"], s1["nnz"] - s0["nnz"], res.iterations, res.status)
        t0 = time.perf_counter()
        for _ in range(3):
            api.set_frame(h, 1, b, db)
            api.slot_size(h, 1)
        tsel = (time.perf_counter() - t0) / 3
        api.destroy(h)
        dt, ev, nnz, it, st = best
        flops = ev * FLOP_PER_EVAL + nnz * FLOP_PER_NNZ_ITER
        out[name] = dict(workload=note, points=list(n), iterations=it, status=st, align_ms=dt * 1e3,
                         us_per_iteration=dt * 1e6 / max(it, 1), set_frame_ms=tsel * 1e3, evals=ev, nnz_sum=nnz,
                         evals_per_s=ev / dt, frac_of_eval_roofline=ev / dt / eval_roof,
                         tflops=flops / dt / 1e12, frac=flops / dt / fp32_peak)

    tum = capi.TUM1_CALIB()
    a, da, b, db, _ = synth.make_pair(1, tum)
    leg("c1_single_pair", tum, a, da, b, db, api.default_params(), 5,
        "C1: 640x480 pair, TUM fr1 intrinsics, 1 deg / 2.7 cm, defaults; handle path (cluster of 16 CTAs)")
    a, da, b, db, _ = synth.make_pair(3, tum, high_gradient=True, rot_deg=0.8, trans=(0.015, -0.01, 0.012))
    p = api.default_params()
    p.num_want = 60000
    leg("c3_dense_pair", tum, a, da, b, db, p, 3,
        "C3: dense selection stress (num_want 60000 -> pot 1), one pair on a cooperative grid of 128 CTAs")
    eth = capi.ETH3D_CALIB()
    t = np.array([0.10, -0.05, 0.10])
    t = t / np.linalg.norm(t) * 0.15
    a, da, b, db, _ = synth.make_pair(4, eth, w=739, h=458, rot_deg=8.0, trans=tuple(t))
    leg("c4_eth3d_pair", eth, a, da, b, db, api.default_params(), 2,
        "C4: 739x458 pair, ETH3D intrinsics, 8 deg / 0.15 m, defaults (the schedule leaves the basin: ~740 iterations)")
    p = api.default_params()
    p.ell_init = 0.25
    leg("c4_eth3d_pair_ell025", eth, a, da, b, db, p, 2, "C4 with ell_init = 0.25 (wide cutoff)")
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_frames, n_pairs = a.frames, a.frames * a.partners
    seed = 1000
    poses = keyframe_poses(n_frames, seed)
    pairs = pair_list(n_frames, a.partners)
    R0, T0, gts = pair_priors(pairs, poses, seed)
    strong = a.scaling == "strong"
    if strong:
        wl = (f"C5 batch: ONE list of {n_pairs} independent pairs over {n_frames} synthetic 640x480 keyframes "
              f"({a.partners} partners each, loop-closure verification shape), TUM fr1 intrinsics, prior via reset_initial, "
              f"sharded over the ranks in contiguous blocks; step = select the block's frames + align its pairs + inner products")
    else:
        wl = (f"C5 batch: {n_frames} synthetic 640x480 keyframes x {a.partners} partners = {n_pairs} independent pairs "
              f"per GPU (loop-closure verification shape), TUM fr1 intrinsics, prior via reset_initial; "
              f"step = select {n_frames} frames + align {n_pairs} pairs + inner products")
    config = dict(workload=wl, frames=n_frames, pairs=n_pairs, pairs_per_gpu=n_pairs // world if strong else n_pairs,
                  exp_mode=a.exp_mode, sharding="contiguous blocks of pairs, no data-path collective, final gather" if strong else "replicas",
                  l2="inputs per step (%.0f MB per GPU at N=1) exceed the 126 MB L2" % (n_frames * W * H * 5 / 1e6))

    if a.impl == "reference":
        if rank != 0:
            return
        import torch
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        cb, ms = cpu_leg(pairs, poses, R0, T0, seed, a.cpu_pairs, a.steps, a.warmup, dev, a.partners)
        line = dict(metric="cvo_frame_pair_alignments_per_s", value=cb["value"], unit="alignments/s", n_gpus=a.gpus,
                    steps=a.steps, warmup=a.warmup, ms_per_step=ms, higher_is_better=True, scaling=a.scaling,
                    vs_baseline=None, dtype="f32 (f64 exp)", data="synthetic", config=config, impl="reference",
                    cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit="alignments/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from cvo_slam_b200 import batch as B, capi, parallel
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    api = capi.load()
    cal = capi.TUM1_CALIB()
    prm = api.default_params()
    prm.exp_mode = a.exp_mode
    pairs_np = np.asarray(pairs, dtype=np.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_arm(idx, noise_salt, steps, warmup, exp_mode, with_e2e=True, sample_clocks=False, collective=True):
        """Times `steps` steps over the pairs `idx` (global indices) on this rank -> dict of measurements.
        collective=False: an arm only this rank runs (no barrier across ranks)."""
        def barrier():
            if world > 1 and collective:
                dist.barrier()
            torch.cuda.synchronize()

        need = parallel.frames_needed(pairs_np, idx)
        local = {int(f): k for k, f in enumerate(need)}
        lp = [(local[int(f)], local[int(m)]) for f, m in pairs_np[idx]]
        bgr_d, dep_d = render_frames([int(f) for f in need], poses, seed, dev, noise_salt=noise_salt)
        torch.cuda.synchronize()
        pr = api.default_params()
        pr.exp_mode = exp_mode
        nf = len(need)
        # the arena holds two ranges of nf frames: the end-to-end leg uploads the frames of step k+1 into one range
        # while step k is aligned out of the other (the device-resident leg only uses the first)
        bt = B.Batch(cal, pr, max_frames=2 * nf, max_pairs=max(1, len(idx)), width=W, height=H, device=local_rank, api=api)
        desc = bt.make_pairs(lp, R0[idx].reshape(-1, 3, 3), T0[idx], pr.ell_init)
        desc_bank = [desc, bt.make_pairs([(f + nf, m + nf) for f, m in lp], R0[idx].reshape(-1, 3, 3), T0[idx], pr.ell_init)]

        def step_device():
            bt.mark(0)
            bt.set_frames_ptr(bgr_d.data_ptr(), dep_d.data_ptr(), nf, device=True)
            res = bt.align(desc)
            vals, nums = bt.inner_product(desc, res)
            bt.mark(1)
            return res, vals, bt.elapsed_ms(), bt.last_align_ms()

        for _ in range(warmup):
            res, vals, _, _ = step_device()
        s0 = bt.stats()
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        t_dev = t_align = 0.0
        wall0 = time.perf_counter()
        for _ in range(steps):
            res, vals, ms, ams = step_device()
            t_dev += ms
            t_align += ams
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sampler else None
        s1 = bt.stats()
        out = dict(res=res, vals=vals, ms_step=t_dev / steps, align_ms=t_align / steps, wall_ms=wall / steps * 1e3,
                   clocks=clocks, stats0=s0, stats1=s1, phases=bt.phase_cycles(), frames=nf, desc_bytes=desc.nbytes,
                   idx=idx)
        if with_e2e:
            # end to end: host (pinned) images through the C ABI, H2D of every frame and D2H of every result inside the
            # timed region; with N ranks the final gather of the results on rank 0 is inside as well
            bgr_h = torch.empty(bgr_d.shape, dtype=torch.uint8, pin_memory=True).copy_(bgr_d)
            dep_h = torch.empty(dep_d.shape, dtype=torch.int16, pin_memory=True).copy_(dep_d)
            torch.cuda.synchronize()
            per = (n_pairs + world - 1) // world
            gbuf = torch.zeros(per * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
            glist = [torch.zeros_like(gbuf) for _ in range(world)] if (world > 1 and rank == 0) else None

            # Every step uploads its own nf frames from pinned host memory (H2D + selection) and reads its results
            # back; the upload of step k+1 is enqueued before step k is aligned and goes to the other half of the
            # arena, so it runs beside that alignment (cvo_batch_set_frames only enqueues).  All `steps` uploads,
            # the first one included, are inside the timed region.
            def upload(k):
                bt.set_frames_ptr(bgr_h.data_ptr(), dep_h.data_ptr(), nf, first=(k & 1) * nf, device=False)

            def compute(k):
                d = desc_bank[k & 1]
                r = bt.align(d)
                v, n = bt.inner_product(d, r)
                if world > 1 and strong:
                    raw = torch.from_numpy(r.view(np.uint8).reshape(-1))
                    gbuf[:raw.numel()].copy_(raw, non_blocking=True)
                    dist.gather(gbuf, glist, dst=0)
                return r, v

            def run_host(k_steps):
                rk = vk = None
                if k_steps < 1:
                    return rk, vk
                upload(0)
                for k in range(k_steps):
                    if k + 1 < k_steps:
                        upload(k + 1)
                    rk, vk = compute(k)
                return rk, vk

            run_host(min(warmup, 2))
            barrier()
            e0 = time.perf_counter()
            res_h, vals_h = run_host(steps)
            barrier()
            out["e2e_s"] = (time.perf_counter() - e0) / steps
            assert np.array_equal(res_h["transform"], res["transform"])   # the two legs agree to the bit
            if glist is not None:   # rank 0: the gathered records, in pair order
                full = np.zeros(n_pairs, dtype=capi.RESULT_DTYPE)
                for r_, g in enumerate(glist):
                    ir = parallel.partition_blocks(n_pairs, r_, world)
                    full[ir] = np.frombuffer(g.cpu().numpy().tobytes()[:len(ir) * capi.RESULT_DTYPE.itemsize], dtype=capi.RESULT_DTYPE)
                out["gathered"] = full
        out["bt"] = bt
        return out

    idx = parallel.partition_blocks(n_pairs, rank, world) if strong else np.arange(n_pairs, dtype=np.int64)
    m = run_arm(idx, 0 if strong else rank, a.steps, a.warmup, a.exp_mode, sample_clocks=True)
    res, bt = m["res"], m["bt"]
    print("phase cycles (cumulative):", m["phases"], "stats", m["stats1"], file=sys.stderr)
    err = []
    for k in range(0, len(idx), max(1, len(idx) // 64)):
        E = np.linalg.inv(gts[int(idx[k])]) @ res["transform"][k].reshape(4, 4).astype(np.float64)
        err.append(np.linalg.norm(E[:3, 3]))
    status_bad = int((res["status"] != 0).sum())

    # inside the run: the sharded job reproduces the single-GPU results to the bit.  Rank 0 re-aligns a sample of
    # pairs from every rank's block on its own GPU (the full list at N = 1 IS the single-GPU run).
    shard_check = None
    if strong and world > 1 and rank == 0 and "gathered" in m:
        samp = np.unique(np.linspace(0, n_pairs - 1, 128).astype(np.int64))
        chk = run_arm(samp, 0, 1, 0, a.exp_mode, with_e2e=False, collective=False)
        same = bool(np.array_equal(chk["res"]["transform"], m["gathered"]["transform"][samp])
                    and np.array_equal(chk["res"]["iterations"], m["gathered"]["iterations"][samp]))
        chk["bt"].close()
        shard_check = dict(pairs_rechecked_on_rank0=int(len(samp)), bit_identical=same)
        assert same, "sharded results differ from the single-GPU results"

    ms_step, e2e_s = m["ms_step"], m["e2e_s"]
    rank_ms = [[round(ms_step, 2), round(m["align_ms"], 2), m["frames"], len(idx)]]
    if world > 1:
        mine = torch.tensor([ms_step, m["align_ms"], m["frames"], len(idx)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [[round(float(x[0]), 2), round(float(x[1]), 2), int(x[2]), int(x[3])] for x in allr]
        t = torch.tensor([ms_step, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s = float(t[0]), float(t[1])

    weak_side = None
    if strong and world > 1:   # the round-1 variant beside it: every rank the full list (own sensor noise)
        bt.close()
        wk = run_arm(np.arange(n_pairs, dtype=np.int64), rank, 2, 1, a.exp_mode, with_e2e=False)
        t = torch.tensor([wk["ms_step"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        weak_side = dict(scaling="weak", value=n_pairs * world / (float(t[0]) * 1e-3), unit="alignments/s",
                         ms_per_step=float(t[0]), pairs_per_gpu=n_pairs, steps=2, warmup=1)
        wk["bt"].close()
        bt = None
    if rank != 0:
        if bt is not None:
            bt.close()
        if world > 1:
            dist.destroy_process_group()
        return

    total_pairs = n_pairs if strong else n_pairs * world
    value = total_pairs / (ms_step * 1e-3)
    s0, s1, clocks = m["stats0"], m["stats1"], m["clocks"]
    my_pairs = len(idx)
    evals = (s1["evals"] - s0["evals"]) / a.steps
    nnz_it = (s1["nnz"] - s0["nnz"]) / a.steps
    iters = (s1["iterations"] - s0["iterations"]) / a.steps
    launches = (s1["launches"] - s0["launches"]) // a.steps
    align_ms = m["align_ms"]
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = SM_COUNT * FP32_LANES * 2 * sm_mhz * 1e6 / 1e12

    def roofline_of(evals, nnz_it, iters, align_ms, ms_step, exp_mode, phases, pairs_here, cum_iters):
        flops = evals * FLOP_PER_EVAL + nnz_it * FLOP_PER_NNZ_ITER
        achieved = flops / (align_ms * 1e-3) / 1e12
        traffic, traffic_src = committed_ncu_traffic(n_frames, a.partners, exp_mode) if pairs_here == n_pairs else (None, None)
        hbm_peak, hbm_src = measured_hbm_peak()
        return dict(bound="fp32 (non-tensor; exact mode adds 1 fp64 exp per eval and exact accumulation)" if exp_mode == 0 else "fp32+mufu",
                    kernel="k_align_batch", achieved=achieved, peak=fp32_peak, unit="TFLOP/s",
                    frac=achieved / fp32_peak, traffic=traffic,
                    peak_source=f"derived: {SM_COUNT} SM x {FP32_LANES} lanes x 2 x observed SM clock {sm_mhz:.0f} MHz "
                                f"(MEASURED_PEAKS.json has no fp32 figure)",
                    kernel_ms_per_launch=align_ms, kernel_share_of_step=align_ms / ms_step, pairs_per_launch=pairs_here,
                    evals_per_launch=evals, evals_per_s=evals / (align_ms * 1e-3),
                    eval_roofline_per_s=min(fp32_peak * 1e12 / FLOP_PER_EVAL,
                                            SM_COUNT * MUFU_LANES * sm_mhz * 1e6 / 2),
                    iterations_per_pair=iters / max(pairs_here, 1), nnz_per_iteration=nnz_it / max(iters, 1),
                    traffic_unit="bytes of DRAM read + write per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                    traffic_source=traffic_src,
                    hbm=None if traffic is None else dict(
                        achieved=traffic / (align_ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                        frac=traffic / (align_ms * 1e-3) / 1e9 / hbm_peak, peak_source=hbm_src,
                        note="neighbour / verdict lists streamed per iteration (see DESIGN section 3), not algorithmic bytes: "
                             "the clouds themselves are 0.2 MB per pair"),
                    phase_share={k: round(v / max(1, sum(v2 for k2, v2 in phases.items() if k2 not in ("rebuilds", "filters"))), 4)
                                 for k, v in phases.items() if k not in ("rebuilds", "filters")},
                    list_builds_per_pair=dict(
                        searches=round(phases["rebuilds"] / max(1, cum_iters) * iters / max(pairs_here, 1), 2),
                        filters=round(phases["filters"] / max(1, cum_iters) * iters / max(pairs_here, 1), 2),
                        note="neighbour-list constructions per alignment: grid searches, and filters of the current list at a "
                             "length-scale change (DESIGN section 5)"))

    roofline = roofline_of(evals, nnz_it, iters, align_ms, m["ms_step"], a.exp_mode, m["phases"], my_pairs, s1["iterations"])
    if world > 1:
        roofline["note"] = "rank 0's launch (its block of the list)"
    line = dict(metric="cvo_frame_pair_alignments_per_s", value=value, unit="alignments/s", n_gpus=world,
                steps=a.steps, warmup=a.warmup, ms_per_step=ms_step, higher_is_better=True, scaling=a.scaling,
                vs_baseline=None, dtype="f32 (f64 exp)" if a.exp_mode == 0 else "f32", data="synthetic",
                config=config, clocks=clocks,
                e2e=dict(value=total_pairs / e2e_s, unit="alignments/s",
                         h2d_bytes_per_step=int(m["frames"] * W * H * 5 + m["desc_bytes"] * 2),
                         d2h_bytes_per_step=int(res.nbytes + my_pairs * 192),
                         note="bytes of rank 0; the timed region includes the gather of all results on rank 0" if world > 1 else None),
                gpu_launches=int(launches * a.steps), roofline=roofline,
                wall_ms_per_step=m["wall_ms"], per_rank_ms_step_align_frames_pairs=rank_ms,
                check=dict(median_translation_error_m=float(np.median(err)), pairs_with_error_status=status_bad))
    if shard_check:
        line["shard_check"] = shard_check
    if weak_side:
        line["weak"] = weak_side
    if world == 1 and not a.no_cpu_baseline:
        cb, _ = cpu_leg(pairs, poses, R0, T0, seed, a.cpu_pairs, 1, 0, dev, a.partners)
        line["cpu_baseline"] = cb
    if world == 1:
        # side measurement: the rest of a loop-closure verification (cvo::compute_innerproduct_lc,
        # cvo.cpp:505-561, + the accept test of keyframe_graph.cpp:711-712) for every pair of the step,
        # one launch.  lc_prior = the prior the alignment started from; prior / lc_prior_2 = that prior
        # under two small perturbations (stand-ins for the motion-model and PnP-RANSAC estimates).
        desc = bt.make_pairs(pairs, R0.reshape(-1, 3, 3), T0, prm.ell_init)
        Rt = np.tile(np.eye(4, dtype=np.float32), (n_pairs, 1, 1))
        Rt[:, :3, :3] = R0.reshape(-1, 3, 3)
        Rt[:, :3, 3] = T0
        lc_prior = np.linalg.inv(Rt.astype(np.float64)).astype(np.float32)
        rng = np.random.default_rng(seed + 7)
        def jitter(scale):
            d = np.tile(np.eye(4, dtype=np.float32), (n_pairs, 1, 1))
            d[:, :3, 3] = rng.normal(0, scale, (n_pairs, 3))
            return (lc_prior @ d).astype(np.float32)
        prior, lc_prior2 = jitter(5e-3), jitter(2e-3)
        bt.verify_lc(desc, res, prior, lc_prior, lc_prior2)
        torch.cuda.synchronize()
        v0 = time.perf_counter()
        lc = bt.verify_lc(desc, res, prior, lc_prior, lc_prior2)
        v_ms = (time.perf_counter() - v0) * 1e3
        line["lc_verify"] = dict(workload="cvo_batch_verify_lc over the step's pairs (6 queries per pair + self products, host call incl. D2H)",
                                 ms=v_ms, pairs_per_s=n_pairs / (v_ms * 1e-3), accepted_fraction=float(lc["accept"].mean()),
                                 verified_candidates_per_s=n_pairs / ((ms_step + v_ms) * 1e-3))
    if bt is not None:
        bt.close()
    if world == 1 and a.side_legs:
        # the FP32 + MUFU mode of the same launch (the kernel north_star's roofline is defined on), beside the exact headline
        other = 1 - a.exp_mode if a.exp_mode in (0, 1) else 1
        fm = run_arm(np.arange(n_pairs, dtype=np.int64), 0, 2, 2, other, with_e2e=False)
        f0, f1 = fm["stats0"], fm["stats1"]
        key = "roofline_fast" if other == 1 else "roofline_exact"
        line[key] = roofline_of((f1["evals"] - f0["evals"]) / 2, (f1["nnz"] - f0["nnz"]) / 2, (f1["iterations"] - f0["iterations"]) / 2,
                                fm["align_ms"], fm["ms_step"], other, fm["phases"], n_pairs, f1["iterations"])
        line[key]["value_alignments_per_s"] = n_pairs / (fm["ms_step"] * 1e-3)
        fm["bt"].close()
        line["single_pair"] = single_pair_legs(api, sm_mhz)
    if world == 1 and a.sequence_frames >= 8:
        line["sequence_c2"] = sequence_leg(a.sequence_frames, api, dev, 0 if a.no_cpu_baseline else 6)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
