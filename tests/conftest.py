import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Pose tolerances.  BASELINE.json's north_star states 1e-4 rad / 1e-4 m per pair; the parity tests hold the three
# BASELINE configs (1, 2, 3) to exactly that.  The converged pose of this algorithm has an intrinsic numerical noise
# floor of the same order, though: the line search takes the smallest positive root of a cubic (src/cvo.cpp:291-307),
# which jumps discontinuously when two roots merge, and the stop tests fire on 1e-5-sized quantities.  The SAME CPU
# restatement compiled two ways (no FMA contraction + brute-force ball  vs  GCC fp-contract=fast + the reference's
# nanoflann) differs from itself by up to 1.1e-4 rad / 1.7e-4 m (median 2.5e-5 / 3.9e-5) over 24 seeded pairs --
# profiles/oracle_noise_floor_r01.json, scripts/noise_floor.py.  Arbitrary extra seeds are therefore held to
# POSE_TOL_FLOOR = 3e-4, not to 1e-4.
POSE_TOL_NORTH_STAR = 1e-4
POSE_TOL_FLOOR = 3e-4


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


def iters_comparable(a, b):
    """Exit iterations of two correct executions: the stop tests fire on 1e-5-sized quantities (eps, eps_2), so the exit
    iteration is chaotic -- the reference's own align() and its restatement differ by up to 2x on the same pair (59 vs
    45, tests/golden) while their poses agree to 1e-5.  Bounded here to a factor of two plus a slack of 15."""
    a, b = int(a), int(b)
    return abs(a - b) <= max(15, max(a, b) // 2)


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
