"""CPU: the C-ABI library builds, loads and exports every symbol that
include/quadx_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    from fpv_drone_rl_agent_b200 import _lib

    return _lib


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:qx|ppo)_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = _declared("quadx_b200.h")
    assert len(names) >= 15
    L = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/quadx_b200.h but not exported"
    assert sorted(built.EXPORTED) == names
    pnames = _declared("ppo_b200.h")
    for n in pnames:
        assert hasattr(L, n), f"{n} declared in include/ppo_b200.h but not exported"
    assert sorted(built.PPO_EXPORTED) == pnames


def test_config_layout_and_defaults(built):
    from oracle.hover_oracle import HoverConfig
    from oracle.quadx_model import QuadXParams

    L = built.lib()
    assert L.qx_version() == 2
    assert L.qx_sizeof_config() == ctypes.sizeof(built.QxConfig)
    c = built.default_config(built.QX_TASK_HOVER)
    p, h = QuadXParams(), HoverConfig()
    f = lambda x: pytest.approx(x, rel=1e-6)  # noqa: E731
    assert c.mass == f(p.mass) and list(c.inertia) == f(list(p.inertia))
    assert list(c.motor_x) == f([m[0] for m in p.motor_xy]) and list(c.motor_y) == f([m[1] for m in p.motor_xy])
    assert list(c.motor_map) == f([v for row in p.motor_map for v in row]) and list(c.torque_sign) == f(list(p.torque_sign))
    assert c.total_thrust == f(p.total_thrust) and c.thrust_coef == f(p.thrust_coef) and c.torque_coef == f(p.torque_coef)
    assert c.noise_ratio == f(p.noise_ratio) and c.tau == f(p.tau)
    assert c.drag_coef_xyz == f(p.drag_coef_xyz) and c.drag_area_xyz == f(p.drag_area_xyz) and c.drag_coef_pqr == f(p.drag_coef_pqr)
    assert list(c.rate_kp) == f(list(p.rate_kp)) and list(c.rate_ki) == f(list(p.rate_ki)) and list(c.rate_kd) == f(list(p.rate_kd))
    assert c.physics_hz == p.physics_hz and c.control_hz == p.control_hz and c.gravity == f(p.gravity)
    assert bool(c.state_stale) == p.state_stale and bool(c.gyro) == p.gyro and c.floor_z == f(p.floor_z)
    assert c.cam_tilt_up_deg == p.cam_tilt_up_deg and c.cam_fov_deg == p.cam_fov_deg and c.cam_res == p.cam_res
    from oracle import vision

    assert list(c.panel) == f(list(vision.panel_front_face().ravel())) and c.vis_margin_px == vision.VIS_MARGIN_PX
    assert c.aviary_steps_per_step == h.env_step_ratio and c.max_steps == h.max_steps and c.reset_idle_steps == h.reset_idle_steps
    assert c.floor_grace_steps == h.floor_grace_steps and c.agent_dt == f(h.agent_dt)
    assert c.flight_dome_size == h.flight_dome_size and c.floor_threshold == f(h.floor_threshold)
    assert c.target_area == f(h.target_area) and c.target_ratio == f(h.target_ratio) and list(c.action_scale) == list(h.action_scale)
    # flight modes: mode 0 like hover.py:92, (a3 + 1) / 2 thrust mapping, outer-loop gains of cf2x.yaml:21-54
    assert c.flight_mode == p.flight_mode == 0 and c.thrust_scale == h.thrust_scale and c.thrust_bias == h.thrust_bias
    assert list(c.att_pid) == f([*p.att_kp, *p.att_ki, *p.att_kd, *p.att_lim]) and list(c.vel_pid) == f([*p.vel_kp, *p.vel_ki, *p.vel_kd, *p.vel_lim])
    assert list(c.pos_pid) == f([*p.pos_kp, *p.pos_ki, *p.pos_kd, *p.pos_lim])
    assert list(c.zpos_pid) == f(list(p.zpos_pid)) and list(c.zvel_pid) == f(list(p.zvel_pid))


def test_reference_constants_header_is_current(built):
    """csrc/qx_ref_constants.cuh (tools/gen_ref_constants.py) still equals what derive() makes of the default config, so the
    reference's own parameter set takes the specialised kernels; anything else takes the generic ones."""
    import ctypes as C

    L = built.lib()
    c = built.default_config(built.QX_TASK_HOVER)
    assert L.qx_config_matches_reference_constants(C.byref(c)) == 1
    c.update(start_pos=[0, 0, 1.0], spawn_throttle=0.4952, reset_idle_steps=0, max_steps=1000, auto_reset=0, noise=0)  # run-time switches
    assert L.qx_config_matches_reference_constants(C.byref(c)) == 1
    for change in (dict(mass=0.11), dict(rate_kp=[0.03, 0.02, 0.04]), dict(control_hz=240.0), dict(gyro=0), dict(cam_tilt_up_deg=20.0),
                   dict(flight_mode=3), dict(target_area=0.02)):
        c2 = built.default_config(built.QX_TASK_HOVER)
        c2.update(**change)
        assert L.qx_config_matches_reference_constants(C.byref(c2)) == 0, change
    assert L.qx_config_matches_reference_constants(C.byref(built.default_config(built.QX_TASK_YAW))) == 0


def test_no_gpu_means_loud_failure(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fpv_drone_rl_agent_b200 import QuadXSim

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QuadXSim(4)
    # the raw C-ABI refuses too, with an error string
    h = ctypes.c_void_p()
    cfg = built.default_config()
    rc = built.lib().qx_create(ctypes.byref(cfg), 4, 0, 0, 0, ctypes.byref(h))
    assert rc < 0 and built.lib().qx_last_error()


def test_argument_validation_needs_no_gpu(built):
    """Error behaviour of the boundary: bad arguments come back as negative QX_E* codes with
    qx_last_error() set, never as exceptions or crashes -- and are rejected before anything touches CUDA."""
    import ctypes as C

    L = built.lib()
    L.qx_last_error.restype = C.c_char_p
    cfg = built.default_config(built.QX_TASK_HOVER)
    h = C.c_void_p()
    QX_EINVAL = -1
    # empty / negative batches, missing out-parameter, missing config
    for n in (0, -5):
        assert L.qx_create(C.byref(cfg), n, C.c_uint64(0), C.c_uint64(0), 0, C.byref(h)) == QX_EINVAL and not h.value
    assert b"bad arguments" in L.qx_last_error()
    assert L.qx_create(None, 4, C.c_uint64(0), C.c_uint64(0), 0, C.byref(h)) == QX_EINVAL
    assert L.qx_create(C.byref(cfg), 4, C.c_uint64(0), C.c_uint64(0), 0, None) == QX_EINVAL
    assert L.qx_default_config(7, C.byref(cfg)) == QX_EINVAL and L.qx_default_config(0, None) == QX_EINVAL
    # null handles
    assert L.qx_reset(None, None, None, 0, 0, None) == QX_EINVAL
    assert L.qx_step(None, None, None, 0, 0, None, None, None, None, None) == QX_EINVAL
    assert L.qx_step_k(None, 1, None, None, None, None, None, None) == QX_EINVAL
    assert L.qx_step_host(None, None, None, None, None, None, None) == QX_EINVAL
    assert L.qx_get_state(None, None) == QX_EINVAL and L.qx_set_state(None, None) == QX_EINVAL
    assert L.qx_nonfinite_count(None, None) == QX_EINVAL and L.qx_done_queue(None, None, None) == QX_EINVAL
    assert L.qx_destroy(None) == 0 and L.qx_num_envs(None) == 0 and L.qx_obs_dim(None) == 0 and L.qx_state_words(None) == 0
    assert L.qx_uses_reference_constants(None) == 0 and L.qx_config_matches_reference_constants(None) == -1
    # rollout kernels: null buffers, empty batches, too many columns
    assert L.ppo_gae(None, None, None, None, 8, 16, 0.99, 0.95, None, None, None) == QX_EINVAL
    assert L.ppo_running_stats_update(None, 20, 16, 20, None, 1e-8, None, None, None, None) == QX_EINVAL
    assert L.ppo_reward_normalize(None, None, None, None, 16, 0.99, 10.0, 1e-8, None, None, None, None, None) == QX_EINVAL
    assert L.ppo_policy_forward(None, None, 20, 16, None, None, 10.0, 0, 0, 0, None, 0, None, None, None, None, None, None) == QX_EINVAL
    assert b"ppo_policy_forward" in L.qx_last_error()
    assert L.ppo_running_stats_scratch_bytes(20) >= 148 * 64 * 8
