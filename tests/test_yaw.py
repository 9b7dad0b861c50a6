"""The yaw task: CPU checks of the oracle against what yaw.py writes down, and (GPU) the CUDA path against the oracle."""
import numpy as np
import pytest


def test_yaw_oracle_follows_yaw_py():
    from oracle.yaw_oracle import YawVecOracle

    n = 8
    orc = YawVecOracle(n, seed=1, auto_reset=False, noise=False)
    obs = orc.reset()
    assert obs.shape == (n, 12)  # yaw.py:41-45
    assert np.all(obs[:, 8:12] == 0) and np.all(obs[:, 5:8] == 0)  # histories cleared, yaw.py:85-92
    yaw0 = obs[:, 2] * np.pi
    assert np.ptp(yaw0) > 0.5 and np.all(np.abs(yaw0) <= np.pi)  # yaw.py:79 U(-pi, pi)
    a = np.linspace(-1, 1, n)
    hist = np.zeros((n, 4))
    for k in range(6):
        obs, r, te, tr, _ = orc.step(a * (k + 1) / 6)
        hist = np.concatenate([hist[:, 1:], (a * (k + 1) / 6)[:, None]], 1)
        np.testing.assert_array_equal(obs[:, 8:12], hist)  # yaw.py:53-55,108: last 4 yaw actions, newest last
        assert np.all(np.abs(obs) <= 1.0 + 1e-12)  # Box(-1, 1)
    # zero thrust (yaw.py:50,121): the drone stays on the floor and only yaws
    assert np.all(orc.st.pos[:, 2] < 0.05) and np.abs(orc.st.s_euler[:, :2]).max() < 1e-6
    # yaw command sign: setpoint r = -30 a (yaw.py:120); positive a must turn the drone to negative yaw
    orc3 = YawVecOracle(n, seed=2, auto_reset=False, noise=False)
    y0 = orc3.reset()[:, 2] * np.pi
    for k in range(8):
        orc3.step(np.full(n, 0.5))
    assert np.all(((orc3.st.s_euler[:, 2] - y0 + np.pi) % (2 * np.pi) - np.pi) < -0.05)
    # truncation on the step where step_count (pre-increment) reaches max_steps, yaw.py:138-139,145
    orc2 = YawVecOracle(2, seed=0, auto_reset=False, noise=False)
    orc2.reset()
    first = None
    for k in range(405):
        _, _, te, tr, _ = orc2.step(np.zeros(2))
        if tr.any() and first is None:
            first = k + 1
    assert first == 401 and not te.any()


@pytest.mark.gpu
def test_yaw_cuda_matches_oracle():
    import torch

    import __graft_entry__ as ge

    ge.build()
    import fpv_drone_rl_agent_b200 as pkg
    from oracle.yaw_oracle import YawVecOracle

    from oracle.yaw_oracle import YawConfig

    n = 512
    # default camera: PyFlyt's +20 deg = 20 deg DOWN (the convention hover.py's -25 = up implies); from the ground the sphere at
    # 26.6 deg elevation is then above the image (top edge 25 deg), so the parity run uses a camera that does see it
    env0 = pkg.QuadXYawVecEnv(64, seed=4)
    assert float(env0.reset()[:, 3:5].abs().max()) == 0.0 and env0.sim.cfg.cam_tilt_up_deg == -20.0 == YawConfig().cam_tilt_up_deg
    env0.close()
    env = pkg.QuadXYawVecEnv(n, seed=4, cam_tilt_up_deg=20.0)
    orc = YawVecOracle(n, cfg=YawConfig(cam_tilt_up_deg=20.0), seed=4, noise=True)
    o = env.reset().cpu().numpy()
    o2 = orc.reset()
    assert np.abs(o - o2).max() < 2e-3
    rng = np.random.default_rng(0)
    a = (0.3 * rng.uniform(-1, 1, n)).astype(np.float32)
    seen_vis = 0
    for k in range(450):
        if k % 40 == 0:
            a = (0.3 * rng.uniform(-1, 1, n)).astype(np.float32)  # small commands: the drone stays in floor contact
        o, r, d, _ = env.step(torch.as_tensor(a))
        o2, r2, te2, tr2, _ = orc.step(a.astype(np.float64))
        o, r = o.cpu().numpy(), r.cpu().numpy()
        assert np.array_equal(d.cpu().numpy(), te2 | tr2), k
        # a sphere within fp32 rounding of the image border may flip visibility; everything else must agree
        same = (o[:, 3] != 0) == (o2[:, 3] != 0)
        assert same.mean() > 0.995
        cols = [0, 1, 2, 5, 6, 7, 8, 9, 10, 11]
        assert np.abs(o[:, cols] - o2[:, cols]).max() < 3e-3, (k, np.abs(o[:, cols] - o2[:, cols]).max())
        assert np.abs(o[same][:, 3:5] - o2[same][:, 3:5]).max() < 2e-3
        assert np.abs(r[same] - r2[same]).max() < 3e-3
        seen_vis += int((o2[:, 3] != 0).sum())
    assert seen_vis > 1000  # the sphere really enters the image while yawing
    s, l, c = env.sim.episode_stats()
    assert c == n and l == 401 * n  # every env ran into the 401-step time limit exactly once
    env.close()
    # gymnasium-shaped facade
    e1 = pkg.DroneEnv(seed=1)
    ob, info = e1.reset()
    assert ob.shape == (12,) and info == {}
    ob, rew, te, tr, info = e1.step(np.array([0.3], np.float32))
    assert ob.shape == (12,) and ob[11] == np.float32(0.3) and not te and not tr
    e1.close()
