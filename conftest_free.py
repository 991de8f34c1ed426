"""pose error helper shared by bench.py (kept out of tests/ so that bench does not import pytest fixtures)"""
import numpy as np


def pose_err(T_est, T_ref):
    E = np.linalg.inv(np.asarray(T_ref, np.float64)) @ np.asarray(T_est, np.float64)
    ang = float(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))
    return ang, float(np.linalg.norm(E[:3, 3]))
