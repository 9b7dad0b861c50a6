"""A hand-written stabilising controller used ONLY to produce interesting,
in-dome action sequences for golden vectors and parity tests (the reference has
no policy checkpoint: test_hover.py:8-11 loads files that are not committed).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np


def hover_actions(state: np.ndarray, target: np.ndarray, rng: np.random.Generator | None = None, jitter: float = 0.0) -> np.ndarray:
    """state [N,4,3] = Aviary.state rows (body rates, euler, body vel, pos);
    target [N,3] world position.  Returns actions [N,4] in [-1,1] in the
    hover.py:337-341 convention (30, 30, -30 rad/s rates, (a3+1)/2 thrust)."""
    euler, vb, pos = state[:, 1], state[:, 2], state[:, 3]
    cy, sy = np.cos(euler[:, 2]), np.sin(euler[:, 2])
    # world-frame velocity approximated by yaw-rotating the body velocity
    vx = cy * vb[:, 0] - sy * vb[:, 1]
    vy = sy * vb[:, 0] + cy * vb[:, 1]
    err = target - pos
    ax = 2.0 * err[:, 0] - 2.5 * vx
    ay = 2.0 * err[:, 1] - 2.5 * vy
    az = 4.0 * err[:, 2] - 4.0 * vb[:, 2]
    # desired tilt in the yaw-aligned frame
    axb = cy * ax + sy * ay
    ayb = -sy * ax + cy * ay
    pitch_d = np.clip(axb / 9.81, -0.35, 0.35)
    roll_d = np.clip(-ayb / 9.81, -0.35, 0.35)
    p_cmd = 8.0 * (roll_d - euler[:, 0])
    q_cmd = 8.0 * (pitch_d - euler[:, 1])
    r_cmd = 3.0 * (0.0 - euler[:, 2])
    tilt = np.maximum(np.cos(euler[:, 0]) * np.cos(euler[:, 1]), 0.5)
    # thrust force = total_thrust * cmd^2 (pwm -> rpm -> rpm^2), cf2x.yaml:2-3
    thrust = np.sqrt(0.1 * (9.81 + np.clip(az, -4.0, 6.0)) / (4.0 * tilt))
    a = np.stack([p_cmd / 30.0, q_cmd / 30.0, -r_cmd / 30.0, 2.0 * thrust - 1.0], axis=1)
    if rng is not None and jitter > 0.0:
        a = a + jitter * rng.uniform(-1.0, 1.0, a.shape)
    return np.clip(a, -1.0, 1.0)
