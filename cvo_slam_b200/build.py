"""Builds cvo_slam_b200/libcvo_b200.so with nvcc for sm_100a (in-tree, so it travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("CVO_B200_OUT") or os.path.join(HERE, "libcvo_b200.so")
# align.cu is compiled with -fmad=false: every float/double operation of the alignment loop must
# round exactly as the oracle's (no FMA contraction); its hot loops use explicit _rn intrinsics.
SOURCES = {"select.cu": [], "align.cu": ["-fmad=false"], "capi.cu": [], "multi.cu": [], "ingest.cu": []}
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "cvo_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in list(SOURCES) + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CXX", None)   # the image exports a wrapper compiler; let nvcc use the system g++
    env.pop("CC", None)
    objdir = os.environ.get("CVO_B200_OBJDIR") or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src, flags in SOURCES.items():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + flags + list(extra) + os.environ.get("CVO_NVCC_EXTRA", "").split() + \
            ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, env=env)))
        objs.append(obj)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lz"]   # zlib: DEFLATE of the PNG ingest
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link, env=env)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
