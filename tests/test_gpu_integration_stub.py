"""GPU: the ctypes stub printed in INTEGRATION.md section 2, executed as written (raw C-ABI, numpy buffers, no torch in
the call path) -- the documentation must stay runnable."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs():
    import __graft_entry__ as ge

    ge.build()
    from fpv_drone_rl_agent_b200._lib import LIB_PATH, QxConfig

    L = C.CDLL(LIB_PATH)
    L.qx_last_error.restype = C.c_char_p
    L.qx_sizeof_config.restype = C.c_int64
    assert L.qx_sizeof_config() == C.sizeof(QxConfig)

    def make(n_envs, seed=0, device=0, **overrides):
        cfg = QxConfig()
        L.qx_default_config(0, C.byref(cfg))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        h = C.c_void_p()
        rc = L.qx_create(C.byref(cfg), C.c_int64(n_envs), C.c_uint64(seed), C.c_uint64(0), device, C.byref(h))
        if rc:
            raise RuntimeError(L.qx_last_error().decode())
        return h

    def reset(h, n):
        obs = np.empty((n, 20), np.float32)
        assert L.qx_reset_host(h, None, obs.ctypes.data_as(C.c_void_p)) == 0
        return obs

    def step(h, actions):
        n = actions.shape[0]
        obs = np.empty((n, 20), np.float32); rew = np.empty(n, np.float32)
        te = np.empty(n, np.uint8); tr = np.empty(n, np.uint8)
        a = np.ascontiguousarray(actions, np.float32)
        rc = L.qx_step_host(h, a.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p),
                            te.ctypes.data_as(C.c_void_p), tr.ctypes.data_as(C.c_void_p), None)
        if rc:
            raise RuntimeError(L.qx_last_error().decode())
        return obs, rew, te.astype(bool), tr.astype(bool)

    n = 16
    h = make(n, seed=3)
    obs = reset(h, n)
    assert obs.shape == (n, 20) and np.all(obs[:, 13] == 1.0)  # the target panel is in view from the spawn pose
    a = np.zeros((n, 4), np.float32); a[:, 3] = -1.0
    lengths = None
    for k in range(33):
        obs, rew, te, tr = step(h, a)
        if te.any():
            lengths = k + 1
            break
    assert lengths == 32 and te.all() and not tr.any() and np.all(rew < -90)  # floor rule, hover.py:283-290
    # error path: bad arguments give a code and a message, never an exception across the boundary
    assert L.qx_step_host(h, None, None, None, None, None, None) < 0 and b"qx_step_host" in L.qx_last_error()
    assert L.qx_destroy(h) == 0
