import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rot_angle(Ra, Rb):
    """Small-angle-accurate rotation distance (arccos of the trace has ~3e-4 rad resolution in f32)."""
    D = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    S = (D - D.T) / 2
    s = np.linalg.norm([S[2, 1], S[0, 2], S[1, 0]])
    c = (np.trace(D) - 1) / 2
    return float(np.arctan2(s, c))


def pose_diff(Ta, Tb):
    """(rotation angle [rad], translation distance [m]) between two 4x4 poses."""
    Ta, Tb = np.asarray(Ta, np.float64), np.asarray(Tb, np.float64)
    return rot_angle(Ta[:3, :3], Tb[:3, :3]), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def oracle():
    from oracle import cvo_oracle
    cvo_oracle.load("port")
    return cvo_oracle


@pytest.fixture(scope="session")
def gpu_ctx():
    from cvo_rgbd_b200 import capi
    ctx = capi.Context(0, max_points=10240, max_slots=64)
    yield ctx
    ctx.close()
