"""CPU, build container only: re-run the reference's own hover.py (imported
unmodified from /root/reference) and check (1) the committed golden vectors
are what it produces today, (2) the restated detect_rectangle equals the
reference's on random frames.  Skipped where /root/reference does not exist
(the GPU box)."""
import os

import numpy as np
import pytest

REF = "/root/reference/simulation/hover.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def hover():
    from oracle import aviary_facade as af

    return af.import_reference_hover("/root/reference")


def test_reference_constants(hover):
    from oracle.hover_oracle import HoverConfig

    env = hover.QuadXHoverEnv()
    cfg = HoverConfig()
    assert env.env_step_ratio == cfg.env_step_ratio == 6
    assert env.agent_dt == cfg.agent_dt == 0.025
    assert env.max_steps == cfg.max_steps and env.flight_dome_size == cfg.flight_dome_size
    assert env.floor_threshold == cfg.floor_threshold
    assert env.target_area == cfg.target_area and env.target_ratio == cfg.target_ratio
    assert env.action_space.shape == (4,) and env.observation_space.shape == (20,)


def test_drone_parameter_fixture_is_current(golden_dir):
    """tests/golden/cf2x_params.json is what the reference's cf2x.yaml / cf2x.urdf say today."""
    import importlib.util
    import json

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert mg.drone_params("/root/reference") == json.load(open(os.path.join(golden_dir, "cf2x_params.json")))


@pytest.mark.parametrize("name", ["fly_noisy", "floor", "dome", "render_idle", "agent_hz60", "wild_actions"])
def test_golden_vectors_are_current(hover, golden_dir, name):
    """Re-run the reference's hover.py on the recorded actions: the committed fixtures are what it produces today."""
    from oracle import aviary_facade as af
    from oracle.quadx_model import NoiseSource, QuadXParams

    g = np.load(os.path.join(golden_dir, f"hover_ref_{name}.npz"))
    agent_hz = int(g["agent_hz"]) if "agent_hz" in g.files else 40
    render = bool(g["render"]) if "render" in g.files else False
    src = NoiseSource(int(g["seed"]), np.array([int(g["env_id"])], np.uint64), enabled=bool(g["noise"]))
    af.NOISE_CONTEXT.update(source=src, rng_ctr=0, params=QuadXParams(), idle_steps=10, ratio=int(240 / agent_hz))
    env = hover.QuadXHoverEnv(agent_hz=agent_hz, render=render)
    ep = -1
    for k in range(g["actions"].shape[0]):
        if g["episode_start"][k]:
            ep += 1
            obs, _ = env.reset()
            np.testing.assert_array_equal(obs, g["reset_obs"][ep])
        o, r, te, tr, _ = env.step(g["actions"][k].copy())
        af.NOISE_CONTEXT["rng_ctr"] += 1
        np.testing.assert_array_equal(o, g["obs"][k])
        assert r == g["reward"][k] and te == g["terminated"][k] and tr == g["truncated"][k]


def test_detect_rectangle_matches_reference(hover):
    import cv2

    from oracle.hover_oracle import detect_rectangle

    env = hover.QuadXHoverEnv()
    rng = np.random.default_rng(0)
    n_vis = 0
    for i in range(200):
        img = np.full((128, 128, 4), 200, np.uint8)
        pts = rng.uniform(5, 123, (4, 2)) if i % 3 else rng.uniform(-10, 138, (4, 2))
        hull = cv2.convexHull(pts.astype(np.float32)).astype(np.int32)
        cv2.fillConvexPoly(img, hull, (220, 0, 0, 255))
        if i % 7 == 0:  # a second, smaller blob
            cv2.circle(img, (int(rng.uniform(10, 118)), int(rng.uniform(10, 118))), 3, (150, 0, 0, 255), -1)
        ref = env.detect_rectangle(img)
        mine = detect_rectangle(img)
        assert ref[0] == mine[0]
        np.testing.assert_array_equal(ref[1], mine[1])
        assert ref[2] == mine[2] and ref[3] == mine[3]
        n_vis += bool(ref[0])
    assert 20 < n_vis < 200
