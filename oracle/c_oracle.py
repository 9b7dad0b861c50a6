"""ctypes wrapper of oracle/_ref/libquadx_oracle.so (oracle/quadx_oracle.c).
TEST INFRASTRUCTURE ONLY: CPU baseline for bench.py and a fast batched checker."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import vision
from .hover_oracle import HoverConfig
from .quadx_model import QuadXParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libquadx_oracle.so")
f64, i32 = C.c_double, C.c_int32


class OrcConfig(C.Structure):
    _fields_ = [
        ("total_thrust", f64), ("thrust_coef", f64), ("torque_coef", f64), ("noise_ratio", f64), ("tau", f64),
        ("drag_coef_xyz", f64), ("drag_area_xyz", f64), ("drag_coef_pqr", f64), ("air_density", f64),
        ("kp", f64 * 3), ("ki", f64 * 3), ("kd", f64 * 3), ("lim", f64 * 3),
        ("mass", f64), ("inertia", f64 * 3), ("motor_x", f64 * 4), ("motor_y", f64 * 4), ("torque_sign", f64 * 4),
        ("motor_map", f64 * 16), ("pwm_idle", f64), ("physics_hz", f64), ("control_hz", f64), ("gravity", f64),
        ("state_stale", i32), ("gyro", i32), ("max_coord_vel", f64), ("floor_z", f64),
        ("cam_tilt_up_deg", f64), ("cam_fov_deg", f64), ("cam_res", f64), ("cam_near", f64), ("cam_offset", f64 * 3),
        ("vis_margin_px", f64), ("panel", f64 * 12),
        ("env_step_ratio", i32), ("max_steps", i32), ("floor_grace_steps", i32), ("reset_idle_steps", i32),
        ("agent_dt", f64), ("flight_dome_size", f64), ("floor_threshold", f64), ("target_area", f64), ("target_ratio", f64),
        ("action_scale", f64 * 3), ("start_pos", f64 * 3), ("start_rpy", f64 * 3), ("spawn_throttle", f64),
        ("spawn_pos_noise", f64), ("spawn_yaw_noise", f64), ("render", i32), ("auto_reset", i32), ("noise", i32),
        ("flight_mode", i32), ("thrust_scale", f64), ("thrust_bias", f64),
        ("att", f64 * 12), ("vel", f64 * 8), ("lpos", f64 * 8), ("zpos", f64 * 4), ("zvel", f64 * 4),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "quadx_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB_PATH


def _arr(ct, vals):
    return ct(*[float(v) for v in vals])


def make_config(p: QuadXParams, h: HoverConfig, auto_reset: bool, noise: bool) -> OrcConfig:
    c = OrcConfig()
    for k in ("total_thrust", "thrust_coef", "torque_coef", "noise_ratio", "tau", "drag_coef_xyz", "drag_area_xyz", "drag_coef_pqr",
              "air_density", "mass", "pwm_idle", "physics_hz", "control_hz", "gravity", "max_coord_vel", "floor_z",
              "cam_tilt_up_deg", "cam_fov_deg", "cam_near"):
        setattr(c, k, float(getattr(p, k)))
    c.kp, c.ki, c.kd, c.lim = _arr(f64 * 3, p.rate_kp), _arr(f64 * 3, p.rate_ki), _arr(f64 * 3, p.rate_kd), _arr(f64 * 3, p.rate_lim)
    c.inertia = _arr(f64 * 3, p.inertia)
    c.motor_x, c.motor_y = _arr(f64 * 4, [m[0] for m in p.motor_xy]), _arr(f64 * 4, [m[1] for m in p.motor_xy])
    c.torque_sign = _arr(f64 * 4, p.torque_sign)
    c.motor_map = _arr(f64 * 16, [v for row in p.motor_map for v in row])
    c.state_stale, c.gyro = int(p.state_stale), int(p.gyro)
    c.cam_res, c.cam_offset, c.vis_margin_px = float(p.cam_res), _arr(f64 * 3, p.cam_offset), float(vision.VIS_MARGIN_PX)
    c.panel = _arr(f64 * 12, vision.panel_front_face().ravel())
    c.env_step_ratio, c.max_steps, c.floor_grace_steps, c.reset_idle_steps = h.env_step_ratio, h.max_steps, h.floor_grace_steps, h.reset_idle_steps
    c.agent_dt, c.flight_dome_size, c.floor_threshold = h.agent_dt, h.flight_dome_size, h.floor_threshold
    c.target_area, c.target_ratio, c.action_scale = h.target_area, h.target_ratio, _arr(f64 * 3, h.action_scale)
    c.start_pos, c.start_rpy = _arr(f64 * 3, h.start_pos), _arr(f64 * 3, h.start_rpy)
    c.spawn_throttle, c.spawn_pos_noise, c.spawn_yaw_noise = h.spawn_throttle, h.spawn_pos_noise, h.spawn_yaw_noise
    c.render, c.auto_reset, c.noise = int(h.render), int(auto_reset), int(noise)
    c.flight_mode, c.thrust_scale, c.thrust_bias = int(p.flight_mode), h.thrust_scale, h.thrust_bias
    c.att = _arr(f64 * 12, [*p.att_kp, *p.att_ki, *p.att_kd, *p.att_lim])
    c.vel = _arr(f64 * 8, [*p.vel_kp, *p.vel_ki, *p.vel_kd, *p.vel_lim])
    c.lpos = _arr(f64 * 8, [*p.pos_kp, *p.pos_ki, *p.pos_kd, *p.pos_lim])
    c.zpos, c.zvel = _arr(f64 * 4, p.zpos_pid), _arr(f64 * 4, p.zvel_pid)
    return c


class COracle:
    def __init__(self, n_envs: int, seed: int = 0, env_id0: int = 0, params: QuadXParams | None = None, cfg: HoverConfig | None = None,
                 auto_reset: bool = True, noise: bool = True, **hover_overrides):
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_int64, C.c_uint64, C.c_uint64]
        L.orc_sizeof_config.restype = C.c_int64
        for fn in ("orc_destroy", "orc_reset", "orc_step", "orc_stats", "orc_get_state"):
            getattr(L, fn).restype = None
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_stats.argtypes = [C.c_void_p, C.POINTER(f64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_void_p]
        assert L.orc_sizeof_config() == C.sizeof(OrcConfig)
        self.L, self.n = L, n_envs
        self.threads = int(L.orc_max_threads())
        h = cfg or HoverConfig(**hover_overrides)
        self.c = make_config(params or QuadXParams(), h, auto_reset, noise)
        self.h = L.orc_create(C.byref(self.c), n_envs, seed, env_id0)
        self.obs = np.zeros((n_envs, 20))
        self.rew = np.zeros(n_envs)
        self.te = np.zeros(n_envs, np.uint8)
        self.tr = np.zeros(n_envs, np.uint8)
        self.tobs = np.zeros((n_envs, 20))

    def reset(self, mask=None) -> np.ndarray:
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.L.orc_reset(self.h, None if m is None else m.ctypes.data, self.obs.ctypes.data)
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float64).reshape(self.n, 4)
        self.L.orc_step(self.h, a.ctypes.data, self.obs.ctypes.data, self.rew.ctypes.data, self.te.ctypes.data, self.tr.ctypes.data, self.tobs.ctypes.data)
        return self.obs, self.rew, self.te.astype(bool), self.tr.astype(bool), {"terminal_obs": self.tobs}

    def stats(self):
        s, l, n = f64(), C.c_int64(), C.c_int64()
        self.L.orc_stats(self.h, C.byref(s), C.byref(l), C.byref(n))
        return s.value, l.value, n.value

    def state(self) -> np.ndarray:
        out = np.zeros((self.n, 17))
        self.L.orc_get_state(self.h, out.ctypes.data)
        return out

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass
