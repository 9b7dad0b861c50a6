"""CPU: PyFlyt's flight modes -1..7 (the PID cascade of cf2x.yaml:21-54) in the oracles.

hover.py never leaves mode 0 (set_mode(0) at hover.py:92), so nothing in the reference pins these; the checks are
(1) the threaded C restatement equals the numpy one, and (2) the restated cascade is self-consistent: closed loop, each
mode drives the restated rigid body to its setpoint (a sign error anywhere in the chain diverges instead)."""
import numpy as np
import pytest

from oracle.c_oracle import COracle
from oracle.hover_oracle import HoverConfig, HoverVecOracle
from oracle.quadx_model import QuadXParams, QuadXState, aviary_step, spawn
from tests.util import FLIGHT_MODE_SCALING

HOVER_THR = float(np.sqrt(0.1 * 9.81 / 4.0))


@pytest.mark.parametrize("mode", sorted(FLIGHT_MODE_SCALING))
def test_c_oracle_matches_numpy_oracle_in_every_flight_mode(mode):
    n = 32
    rng = np.random.default_rng(mode + 10)
    kw = dict(start_pos=(0.0, 0.0, 1.0), spawn_throttle=HOVER_THR, spawn_pos_noise=0.2, spawn_yaw_noise=1.0, **FLIGHT_MODE_SCALING[mode])
    p = QuadXParams(flight_mode=mode)
    c = COracle(n, seed=3, env_id0=7, params=p, noise=True, **kw)
    o = HoverVecOracle(n, p, HoverConfig(**kw), seed=3, env_id0=7, noise=True)
    np.testing.assert_allclose(c.reset(), o.reset(), rtol=0, atol=1e-9)
    for k in range(60):
        a = rng.uniform(-1, 1, (n, 4))
        if k % 20 >= 10:
            a[:] = a[:1]  # hold a command for a while so that the outer loops settle and integrate
        ob, r, te, tr, _ = c.step(a)
        ob2, r2, te2, tr2, _ = o.step(a)
        assert np.array_equal(te, te2) and np.array_equal(tr, tr2), (mode, k)
        np.testing.assert_allclose(ob, ob2, rtol=0, atol=1e-8)
        np.testing.assert_allclose(r, r2, rtol=0, atol=1e-8)
    st = c.state()
    np.testing.assert_allclose(st[:, 0:3], o.st.pos, rtol=0, atol=1e-9)
    np.testing.assert_allclose(st[:, 13:17], o.st.thr, rtol=0, atol=1e-9)


def _fly(mode, setpoint, aviary_steps, rpy=(0.0, 0.0, 0.0)):
    p = QuadXParams(flight_mode=mode, noise_ratio=0.0)
    st = QuadXState.zeros(1)
    spawn(st, np.ones(1, bool), p, (0.0, 0.0, 1.0), rpy, HOVER_THR)
    sp = np.asarray(setpoint, float)[None]
    sub = 0
    for _ in range(aviary_steps):
        sub = aviary_step(st, sp, p, None, sub, 0, np.zeros(1, np.uint64))
    return st


def test_cascade_is_self_consistent_closed_loop():
    # mode 7: fly to a pose and hold it
    st = _fly(7, (0.8, -0.5, 1.0, 1.6), 2400)
    assert np.allclose(st.pos[0], (0.8, -0.5, 1.6), atol=0.05) and abs(st.s_euler[0, 2] - 1.0) < 0.02
    assert np.abs(st.vel[0]).max() < 0.05 and np.abs(st.s_euler[0, :2]).max() < 0.02
    # mode 6: ground-frame velocity while yawing
    st = _fly(6, (0.5, -0.3, 0.3, 0.0), 240)  # 2 s: the climb-rate loop sags ~0.5 m before its integral holds the weight
    assert np.allclose(st.vel[0, :2], (0.5, -0.3), atol=0.12) and abs(st.s_wb[0, 2] - 0.3) < 0.02
    # modes 4 / 5: body-frame velocity, height / climb rate
    st = _fly(4, (0.5, -0.3, 0.0, 1.5), 1800)
    assert np.allclose(st.s_vb[0, :2], (0.5, -0.3), atol=0.03) and abs(st.pos[0, 2] - 1.5) < 0.1
    st = _fly(5, (0.0, 0.0, 0.0, 0.3), 1200)  # ki = 0.3: the climb-rate loop needs ~10 s
    assert abs(st.vel[0, 2] - 0.3) < 0.05 and np.abs(st.pos[0, :2]).max() < 1e-6
    # modes 1 / 3: attitude hold
    st = _fly(3, (0.1, -0.1, 0.5, 1.2), 840)  # kp = 1: a 1 s time constant
    assert np.allclose(st.s_euler[0], (0.1, -0.1, 0.5), atol=0.01)
    st = _fly(1, (0.0, 0.0, -0.7, 0.0), 840)
    assert abs(st.s_euler[0, 2] + 0.7) < 0.01 and abs(st.vel[0, 2]) < 0.05
    # mode 2: rates + height
    st = _fly(2, (0.0, 0.0, 0.4, 1.4), 1800)
    assert abs(st.s_wb[0, 2] - 0.4) < 0.01 and abs(st.pos[0, 2] - 1.4) < 0.1
    # mode -1: the hover pwm on all four motors holds the drone (no controller in the loop)
    st = _fly(-1, (HOVER_THR,) * 4, 240)
    assert abs(st.pos[0, 2] - 1.0) < 0.05 and np.abs(st.s_euler[0]).max() < 1e-9
