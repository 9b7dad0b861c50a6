"""CPU: the oracle's env layer against the committed golden vectors, which were
produced by the reference's own hover.py (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle.hover_oracle import HoverConfig, HoverVecOracle
from oracle.quadx_model import QuadXParams

SCENARIOS = ["fly_quiet", "fly_noisy", "floor", "dome", "render_idle", "agent_hz60", "wild_actions"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"hover_ref_{name}.npz"))


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_reproduces_reference_hover(golden_dir, name):
    g = _load(golden_dir, name)
    orc = HoverVecOracle(
        1, QuadXParams(), HoverConfig(render=bool(g["render"]) if "render" in g.files else False, agent_hz=int(g["agent_hz"]) if "agent_hz" in g.files else 40), seed=int(g["seed"]), env_id0=int(g["env_id"]),
        auto_reset=False, noise=bool(g["noise"]), vision_mode="raster",
    )
    ep = -1
    for k in range(g["actions"].shape[0]):
        if g["episode_start"][k]:
            ep += 1
            obs = orc.reset()
            np.testing.assert_allclose(obs[0], g["reset_obs"][ep], rtol=0, atol=1e-12)
        obs, r, te, tr, _ = orc.step(g["actions"][k][None])
        np.testing.assert_allclose(obs[0], g["obs"][k], rtol=0, atol=1e-12, err_msg=f"obs step {k}")
        assert abs(r[0] - g["reward"][k]) < 1e-11, k
        assert bool(te[0]) == bool(g["terminated"][k]) and bool(tr[0]) == bool(g["truncated"][k]), k
        np.testing.assert_allclose(orc.st.aviary_state()[0], g["state"][k], rtol=0, atol=1e-12)


def test_episode_length_signature(golden_dir):
    """The reference's TensorBoard runs start at ep_len 32 (floor rule) and
    saturate at 402 (time limit) -- hover.py:275-276,283-290,356."""
    floor = _load(golden_dir, "floor")
    assert int(np.argmax(floor["terminated"])) + 1 == 32
    assert floor["reward"][31] < -99.0 and floor["reward"][30] > -10.0
    fly = _load(golden_dir, "fly_quiet")
    assert int(np.argmax(fly["truncated"])) + 1 == 402
    assert not fly["terminated"].any()
    # sticky flags freeze the physics after the episode ended (hover.py:347-348)
    np.testing.assert_array_equal(fly["state"][402], fly["state"][407])


def test_prev_action_survives_reset(golden_dir):
    """hover.py:31,357: prev_action is not cleared by reset(), so the first
    smoothness penalty of an episode is taken against the previous episode's
    last action."""
    g = _load(golden_dir, "fly_noisy")
    k = int(np.where(g["episode_start"])[0][1])
    a_prev, a = g["actions"][k - 1], g["actions"][k]
    orc = HoverVecOracle(1, seed=int(g["seed"]), env_id0=int(g["env_id"]), auto_reset=False, noise=True, vision_mode="raster")
    orc.reset()
    for j in range(k):
        orc.step(g["actions"][j][None])
    orc.reset()
    _, r, _, _, _ = orc.step(a[None])
    assert abs(r[0] - g["reward"][k]) < 1e-11
    assert np.linalg.norm(a - a_prev) > 1e-3  # the quirk is actually exercised


def test_auto_reset_matches_manual_reset():
    """SB3 VecEnv semantics: on done the returned obs is the first obs of the next
    episode and the terminal obs is reported separately."""
    a = np.tile(np.array([[0.0, 0.0, 0.0, -1.0]]), (3, 1))
    auto = HoverVecOracle(3, seed=5, auto_reset=True, noise=True)
    man = HoverVecOracle(3, seed=5, auto_reset=False, noise=True)
    auto.reset()
    man.reset()
    for k in range(40):
        o1, r1, te1, tr1, info = auto.step(a)
        o2, r2, te2, tr2, _ = man.step(a)
        np.testing.assert_array_equal(r1, r2)
        np.testing.assert_array_equal(te1, te2)
        if te1.any():
            assert k == 31
            np.testing.assert_array_equal(info["terminal_obs"], o2)
            o2 = man.reset()
            np.testing.assert_array_equal(o1, o2)
        else:
            np.testing.assert_array_equal(o1, o2)
    assert auto.n_done == 3 and auto.sum_len == 96


def test_analytic_vision_close_to_rasterised():
    """Bound of the analytic-vs-pixel error of the three vision features over
    random in-dome poses; these are the tolerances the GPU parity
    tests use for obs[7:16]."""
    from oracle import vision
    from oracle.hover_oracle import detect_rectangle
    from oracle.quadx_model import euler_to_quat

    p = QuadXParams()
    rng = np.random.default_rng(1)
    n = 400
    pos = rng.uniform(-2, 2, (n, 3))
    pos[:, 2] = rng.uniform(0.01, 2.5, n)
    rpy = rng.uniform(-0.5, 0.5, (n, 3))
    q = euler_to_quat(rpy)
    v, c, a, r = vision.analytic_features(pos, q, p)
    V = np.zeros(n, bool)
    C = np.zeros((n, 2))
    A = np.zeros(n)
    R = np.zeros(n)
    for i in range(n):
        V[i], C[i], A[i], R[i] = detect_rectangle(vision.render_rgba(pos[i], q[i], p))
    assert (V != v).mean() <= 0.01
    both = V & v
    assert both.sum() > 200
    assert np.abs(C - c)[both].max() * 64 <= 1.5  # pixels
    assert (np.abs(A - a)[both] / A[both]).max() <= 0.10
    assert np.abs(R - r)[both].max() <= 0.35  # one pixel row/column on a ~5 px high quad


def test_drone_parameters_match_reference_files(golden_dir):
    """every number the oracle (and, via tests/test_abi.py, qx_default_config) uses for the drone
    is the one in the reference's cf2x.yaml / cf2x.urdf (fixture written by tests/golden/make_golden.py:drone_params)."""
    import json

    g = json.load(open(os.path.join(golden_dir, "cf2x_params.json")))
    p = QuadXParams()
    for k in ("total_thrust", "thrust_coef", "torque_coef", "noise_ratio", "tau"):
        assert getattr(p, k) == g[f"motor.{k}"], k
    for k in ("drag_coef_xyz", "drag_area_xyz", "drag_coef_pqr"):
        assert getattr(p, k) == g[f"drag.{k}"], k
    for ours, loop in (("rate", "ang_vel"), ("att", "ang_pos"), ("vel", "lin_vel"), ("pos", "lin_pos")):
        for k in ("kp", "ki", "kd", "lim"):
            assert list(getattr(p, f"{ours}_{k}")) == g[f"{loop}.{k}"], (loop, k)
    assert list(p.zpos_pid) == [g[f"z_pos.{k}"][0] for k in ("kp", "ki", "kd", "lim")]
    assert list(p.zvel_pid) == [g[f"z_vel.{k}"][0] for k in ("kp", "ki", "kd", "lim")]
    assert p.mass == g["urdf.mass"] and list(p.inertia) == g["urdf.inertia_diag"]
    assert [list(m) for m in p.motor_xy] == [xyz[:2] for xyz in g["urdf.prop_xyz"]] and all(xyz[2] == 0 for xyz in g["urdf.prop_xyz"])
    assert p.floor_z == g["urdf.collision_box"][2] / 2  # the floor stand-in is the rest height of the collision box
    # the motor map is derived from the prop positions: tau = r x F = (y F, -x F)
    for (x, y), row in zip(p.motor_xy, p.motor_map):
        assert row[0] == np.sign(y) and row[1] == -np.sign(x) and row[3] == 1.0


def test_raster_vision_equals_reference_detector():
    """vision_mode = 1 (oracle/vision.raster_features: row spans of the box silhouette, contour features by Pick's theorem)
    against the reference's real detect_rectangle on rasterised frames, 1000 random in-dome poses: visibility, contour area
    and bounding-box ratio (hover.py:180-213) must be the detector's own numbers; the centre (mean of the approxPolyDP corners,
    hover.py:197-203) within 0.8 px.  These are the tolerances the GPU test of the raster mode uses."""
    from oracle import vision
    from oracle.hover_oracle import detect_rectangle
    from oracle.quadx_model import euler_to_quat

    p = QuadXParams()
    rng = np.random.default_rng(1)
    n = 1000
    pos = rng.uniform(-2, 2, (n, 3))
    pos[:, 2] = rng.uniform(0.01, 2.5, n)
    q = euler_to_quat(rng.uniform(-0.5, 0.5, (n, 3)))
    v, c, a, r = vision.raster_features(pos, q, p)
    V, C, A, R = np.zeros(n, bool), np.zeros((n, 2)), np.zeros(n), np.zeros(n)
    for i in range(n):
        V[i], C[i], A[i], R[i] = detect_rectangle(vision.render_rgba(pos[i], q[i], p))
    assert (V != v).sum() == 0
    both = V & v
    assert both.sum() > 600
    assert np.abs(C - c)[both].max() * 64 <= 0.8  # pixels
    assert (np.abs(A - a)[both] > 1e-12).sum() <= 3 and np.abs(A - a)[both].max() * 128 * 128 <= 2.0  # a boundary pixel in a handful of frames
    assert (np.abs(R - r)[both] > 1e-12).sum() <= 1
    # the reward term built from them (hover.py:296-317)
    rew = lambda cc, aa, rr: -(np.hypot(cc[:, 0], cc[:, 1]) + np.abs(aa - 0.013) + np.abs(rr - 1.53))  # noqa: E731
    assert np.abs(rew(C, A, R) - rew(c, a, r))[both].max() <= 2e-2
