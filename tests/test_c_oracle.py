"""CPU: the threaded C restatement (CPU baseline) against the numpy oracle."""
import numpy as np

from oracle.c_oracle import COracle
from oracle.hover_oracle import HoverConfig, HoverVecOracle


def test_c_oracle_matches_numpy_oracle():
    n = 64
    rng = np.random.default_rng(0)
    for kw in (dict(), dict(start_pos=(0.0, 0.0, 1.0), spawn_throttle=0.4952, spawn_pos_noise=0.2, spawn_yaw_noise=1.0)):
        c = COracle(n, seed=9, env_id0=100, noise=True, **kw)
        p = HoverVecOracle(n, cfg=HoverConfig(**kw), seed=9, env_id0=100, noise=True)
        np.testing.assert_allclose(c.reset(), p.reset(), rtol=0, atol=1e-9)
        for k in range(80):
            a = rng.uniform(-1, 1, (n, 4))
            a[:, :3] *= 0.2
            a[:, 3] = 0.0 + 0.3 * a[:, 3]
            o, r, te, tr, info = c.step(a)
            o2, r2, te2, tr2, info2 = p.step(a)
            assert np.array_equal(te, te2) and np.array_equal(tr, tr2), k
            np.testing.assert_allclose(r, r2, rtol=0, atol=1e-8)
            np.testing.assert_allclose(o, o2, rtol=0, atol=1e-8)
            done = te | tr
            if done.any():
                np.testing.assert_allclose(info["terminal_obs"][done], info2["terminal_obs"][done], rtol=0, atol=1e-8)
        st = c.state()
        np.testing.assert_allclose(st[:, 0:3], p.st.pos, rtol=0, atol=1e-9)
        np.testing.assert_allclose(st[:, 13:17], p.st.thr, rtol=0, atol=1e-9)
        s, l, nd = c.stats()
        assert nd == p.n_done and l == p.sum_len and abs(s - p.sum_ret) < 1e-6 * max(1.0, abs(p.sum_ret))
        assert nd > 0
