"""Batched float64 restatement of the yaw task (/root/reference/simulation/yaw.py).

TEST INFRASTRUCTURE ONLY.  What yaw.py writes down is followed line by line:
spaces (yaw.py:37-45), hard-coded roll / pitch / throttle (:48-50), the 4-deep
yaw action history (:53-55,108), action scaling (:111-122), ONE Aviary.step per
env step (:126-127), truncation / termination (:138-143), observation order
(:57-74).  What the reference calls but never defines (`detect_red_sphere_center`
:63, `calculate_angular_velocity` :66, `add_sphere` :81, `calculate_reward` :147;
its base class `hover.DroneEnv` does not exist, yaw.py:8) is replaced by the
DECLARED stand-ins documented in DESIGN.md -- those are extensions, not parity.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import vision
from .quadx_model import STREAM_SPAWN, STREAM_STEP, NoiseSource, QuadXParams, QuadXState, aviary_step, spawn, u01


@dataclass
class YawConfig:
    max_steps: int = 400
    flight_dome_size: float = 3.0
    agent_dt: float = 1.0 / 120.0  # one Aviary.step (yaw.py:126-127) at control_hz 120
    sphere_pos: tuple = (2.0, 0.0, 1.0)  # main.py:16-23
    cam_tilt_up_deg: float = -20.0  # PyFlyt default camera_angle_degrees = +20; positive = down under the convention hover.py's -25 = 25 deg up implies
    spawn_yaw_noise: float = np.pi  # yaw.py:79
    start_pos: tuple = (0.0, 0.0, 0.0)
    rate_scale: float = 30.0


class YawVecOracle:
    OBS_DIM = 12

    def __init__(self, n_envs, params: QuadXParams | None = None, cfg: YawConfig | None = None, seed=0, env_id0=0, auto_reset=True, noise=True):
        self.n, self.cfg = n_envs, cfg or YawConfig()
        self.p = params or QuadXParams(cam_tilt_up_deg=self.cfg.cam_tilt_up_deg)
        self.auto_reset = auto_reset
        self.env_ids = env_id0 + np.arange(n_envs, dtype=np.uint64)
        self.noise = NoiseSource(seed, self.env_ids, enabled=noise and self.p.noise_ratio != 0.0)
        self.st = QuadXState.zeros(n_envs)
        self.step_count = np.zeros(n_envs, np.int64)
        self.terminated = np.zeros(n_envs, bool)
        self.truncated = np.zeros(n_envs, bool)
        self.hist = np.zeros((n_envs, 4))  # yaw.py:53-55
        self.prev_euler = np.zeros((n_envs, 3))
        self.rng_ctr = np.zeros(n_envs, np.uint64)

    def _sphere(self):
        eye, fwd, right, up = vision.camera_frame(self.st.pos, self.st.quat, self.p)
        px, py, depth = vision.project(np.asarray([self.cfg.sphere_pos]), eye, fwd, right, up, self.p)
        px, py, depth = px[:, 0], py[:, 0], depth[:, 0]
        res = self.p.cam_res
        vis = (depth > self.p.cam_near) & (px >= 0) & (px <= res) & (py >= 0) & (py <= res)
        return vis, np.where(vis, px / (res / 2) - 1.0, 0.0), np.where(vis, py / (res / 2) - 1.0, 0.0)

    def _obs(self):
        e = self.st.s_euler
        d = (e - self.prev_euler + np.pi) % (2 * np.pi) - np.pi
        w = np.clip(d / self.cfg.agent_dt / self.cfg.rate_scale, -1.0, 1.0)
        vis, cx, cy = self._sphere()
        return np.concatenate([e / np.pi, cx[:, None], cy[:, None], w, self.hist], axis=1), vis, cx  # yaw.py:69-74

    def _reset_envs(self, mask):
        if not mask.any():
            return
        n = self.n
        pos = np.broadcast_to(np.asarray(self.cfg.start_pos, float), (n, 3)).copy()
        rpy = np.zeros((n, 3))
        u = u01(self.noise.bits(0, STREAM_SPAWN, self.rng_ctr)) * 2.0 - 1.0
        rpy[:, 2] += self.cfg.spawn_yaw_noise * u[:, 3]  # yaw.py:79 (applied to this reset, not the next)
        spawn(self.st, mask, self.p, pos, rpy, 0.0)
        self.step_count[mask] = 0
        self.terminated[mask] = False
        self.truncated[mask] = False
        self.hist[mask] = 0.0  # yaw.py:90-92
        self.prev_euler[mask] = self.st.s_euler[mask]

    def reset(self, mask=None):
        mask = np.ones(self.n, bool) if mask is None else np.asarray(mask, bool)
        self._reset_envs(mask)
        return self._obs()[0]

    def step(self, actions):
        a = np.asarray(actions, float).reshape(self.n)
        self.hist = np.concatenate([self.hist[:, 1:], a[:, None]], axis=1)  # yaw.py:108
        sp = np.zeros((self.n, 4))
        sp[:, 2] = a * -30.0  # yaw.py:120
        sp[:, 3] = (-1.0 + 1.0) / 2  # yaw.py:121
        aviary_step(self.st, sp, self.p, self.noise.normals, 0, STREAM_STEP, self.rng_ctr)  # yaw.py:126-127
        self.rng_ctr += np.uint64(1)
        self.truncated |= self.step_count >= self.cfg.max_steps  # yaw.py:138-139
        self.terminated |= np.linalg.norm(self.st.s_pos, axis=1) > self.cfg.flight_dome_size  # yaw.py:141-143
        self.step_count = self.step_count + 1  # yaw.py:145
        obs, vis, cx = self._obs()
        reward = np.where(vis, 1.0 - np.abs(cx), -1.0) - 0.05 * np.abs(self.hist[:, 3] - self.hist[:, 2])  # declared stand-in
        self.prev_euler = self.st.s_euler.copy()
        te, tr = self.terminated.copy(), self.truncated.copy()
        info = {"terminal_obs": obs.copy()}
        done = te | tr
        if self.auto_reset and done.any():
            self._reset_envs(done)
            obs = np.where(done[:, None], self._obs()[0], obs)
        return obs, reward, te, tr, info
