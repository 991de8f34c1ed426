import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_api():
    """CPU oracle (test infrastructure).  Uses the reference's nanoflann where built."""
    from oracle import oracle
    return oracle.load()


@pytest.fixture(scope="session")
def oracle_plain():
    from oracle import oracle
    return oracle.load(kd=False)


@pytest.fixture(scope="session")
def cuda_api():
    """libcvo_b200.so through ctypes; raises (not skips) when the library is missing."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cvo_slam_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def tum_calib():
    from cvo_slam_b200 import capi
    return capi.TUM1_CALIB()


@pytest.fixture(scope="session")
def golden_small():
    path = os.path.join(GOLDEN, "pair_640x480_seed11.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def pair_c1(tum_calib):
    """C1: 640x480 synthetic pair, TUM fr1 intrinsics, 1 deg / 2.7 cm offset, seed 1."""
    from cvo_slam_b200 import synth
    return synth.make_pair(1, tum_calib)


def pose_error(T_est, T_ref):
    """(rotation angle in rad, translation distance in m) between two 4x4 transforms."""
    E = np.linalg.inv(np.asarray(T_ref, np.float64)) @ np.asarray(T_est, np.float64)
    ang = float(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))
    return ang, float(np.linalg.norm(E[:3, 3]))
