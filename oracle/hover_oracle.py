"""Batched float64 restatement of ``QuadXHoverEnv`` (/root/reference/simulation/hover.py)
with the SB3 ``VecEnv`` auto-reset semantics ``train_hover.py:42`` relies on.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The env layer here is
pinned against the reference's own hover.py (tests/test_oracle_vs_reference.py);
the drone model underneath (oracle/quadx_model.py) is a restatement, parity
unpinned.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import vision
from .quadx_model import (
    STREAM_RESET,
    STREAM_SPAWN,
    STREAM_STEP,
    NoiseSource,
    QuadXParams,
    QuadXState,
    aviary_step,
    euler_to_quat,
    mode_preset_setpoint,
    spawn,
    u01,
)


@dataclass
class HoverConfig:
    """Literals of QuadXHoverEnv.__init__ / reset (hover.py:23-51, 77-110)."""

    physics_hz: float = 240.0  # hover.py:23
    agent_hz: int = 40  # hover.py:14
    max_steps: int = 400  # hover.py:35
    flight_dome_size: float = 3.0  # hover.py:37
    floor_threshold: float = 0.1  # hover.py:38
    floor_grace_steps: int = 30  # hover.py:283
    target_area: float = 0.013  # hover.py:50
    target_ratio: float = 1.53  # hover.py:51
    action_scale: tuple = (30.0, 30.0, -30.0)  # hover.py:338-340
    thrust_scale: float = 0.5  # hover.py:341: (a3 + 1) / 2 == a3 * 0.5 + 0.5
    thrust_bias: float = 0.5
    start_pos: tuple = (0.0, 0.0, 0.0)  # hover.py:78
    start_rpy: tuple = (0.0, 0.0, 0.0)  # hover.py:79
    reset_idle_steps: int = 10  # hover.py:109
    # additions (SURVEY 8b "semantic deltas"): all zero = reference behaviour
    spawn_throttle: float = 0.0
    spawn_pos_noise: float = 0.0  # uniform +- on x, y, z
    spawn_yaw_noise: float = 0.0  # uniform +- on yaw
    render: bool = False  # hover.py:283 floor rule is off when rendering

    @property
    def env_step_ratio(self) -> int:
        return int(self.physics_hz / self.agent_hz)  # hover.py:24

    @property
    def agent_dt(self) -> float:
        return 1.0 / self.agent_hz  # hover.py:25


def detect_rectangle(rgba_image: np.ndarray):
    """Restatement of hover.py:157-222 (red mask, border reject, largest
    external contour, 4-vertex approxPolyDP, centre / area / bbox ratio)."""
    import cv2

    h, w = rgba_image.shape[:2]
    red = ((rgba_image[:, :, 0] > 100) & (rgba_image[:, :, 1] == 0) & (rgba_image[:, :, 2] == 0)).astype(np.uint8) * 255
    at_edges = red[0, :].any() or red[-1, :].any() or red[:, 0].any() or red[:, -1].any()
    contours, _ = cv2.findContours(red, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if contours and not at_edges:
        contour = max(contours, key=cv2.contourArea)
        approx = cv2.approxPolyDP(contour, 0.04 * cv2.arcLength(contour, True), True)
        if len(approx) == 4:
            c = approx.reshape((4, 2))
            centre = np.array([np.mean(c[:, 0]) / (w / 2.0) - 1.0, np.mean(c[:, 1]) / (h / 2.0) - 1.0], dtype=np.float32)
            area = cv2.contourArea(contour) / (w * h)
            _, _, bw, bh = cv2.boundingRect(contour)
            return True, centre, area, (bw / bh if bh > 0 else 0.0)
    return False, np.zeros(2, np.float32), 0.0, 0.0


class HoverVecOracle:
    """N independent QuadXHoverEnv instances stepped in lock-step."""

    OBS_DIM = 20

    def __init__(
        self,
        n_envs: int,
        params: QuadXParams | None = None,
        cfg: HoverConfig | None = None,
        seed: int = 0,
        env_id0: int = 0,
        auto_reset: bool = True,
        noise: bool = True,
        vision_mode: str = "analytic",
    ):
        self.n = n_envs
        self.p = params or QuadXParams()
        self.cfg = cfg or HoverConfig()
        self.auto_reset = auto_reset
        self.vision_mode = vision_mode
        self.env_ids = env_id0 + np.arange(n_envs, dtype=np.uint64)
        self.noise = NoiseSource(seed, self.env_ids, enabled=noise and self.p.noise_ratio != 0.0)
        self.st = QuadXState.zeros(n_envs)
        n = n_envs
        self.step_count = np.zeros(n, np.int64)
        self.terminated = np.zeros(n, bool)
        self.truncated = np.zeros(n, bool)
        self.action = np.zeros((n, 4))
        self.prev_action = np.zeros((n, 4))  # hover.py:31 -- NOT cleared by reset()
        self.prev_centre = np.zeros((n, 2))
        self.prev_area = np.zeros(n)
        self.prev_ratio = np.zeros(n)
        self.prev_euler = np.zeros((n, 3))
        self.rng_ctr = np.zeros(n, np.uint64)
        self.ep_return = np.zeros(n)
        self.info = {"out_of_bounds": np.zeros(n, bool), "on_floor": np.zeros(n, bool)}
        self.sum_ret = 0.0
        self.sum_len = 0
        self.n_done = 0

    # ------------------------------------------------------------------ vision
    def _vision(self, mask=None):
        if self.vision_mode == "analytic":
            return vision.analytic_features(self.st.pos, self.st.quat, self.p)
        vis = np.zeros(self.n, bool)
        cen = np.zeros((self.n, 2))
        area = np.zeros(self.n)
        ratio = np.zeros(self.n)
        for i in range(self.n):
            img = vision.render_rgba(self.st.pos[i], self.st.quat[i], self.p)
            vis[i], cen[i], area[i], ratio[i] = detect_rectangle(img)
        return vis, cen, area, ratio

    # ------------------------------------------------------------- compute_state
    def _compute_state(self) -> np.ndarray:
        """hover.py:224-272."""
        euler = self.st.s_euler
        diff = euler - self.prev_euler
        diff = (diff + np.pi) % (2 * np.pi) - np.pi  # hover.py:229
        ang_vel = diff / self.cfg.agent_dt  # hover.py:230
        quat = euler_to_quat(euler)  # hover.py:233
        vis, cen, area, ratio = self._vision()
        obs = np.concatenate(
            [
                ang_vel,
                quat,
                cen,
                self.prev_centre,
                area[:, None],
                self.prev_area[:, None],
                vis.astype(float)[:, None],
                ratio[:, None],
                self.prev_ratio[:, None],
                self.action,
            ],
            axis=1,
        )  # hover.py:253-267
        self.prev_centre = cen.copy()  # hover.py:270-272
        self.prev_area = area.copy()
        self.prev_ratio = ratio.copy()
        return obs

    # ------------------------------------------------------------------- reset
    def _reset_envs(self, mask: np.ndarray) -> None:
        """hover.py:72-113 for the envs in ``mask`` (everything but compute_state)."""
        if not mask.any():
            return
        cfg, p = self.cfg, self.p
        n = self.n
        pos = np.broadcast_to(np.asarray(cfg.start_pos, float), (n, 3)).copy()
        rpy = np.broadcast_to(np.asarray(cfg.start_rpy, float), (n, 3)).copy()
        if cfg.spawn_pos_noise != 0.0 or cfg.spawn_yaw_noise != 0.0:
            u = u01(self.noise.bits(0, STREAM_SPAWN, self.rng_ctr)) * 2.0 - 1.0
            pos += cfg.spawn_pos_noise * u[:, :3]
            rpy[:, 2] += cfg.spawn_yaw_noise * u[:, 3]
        spawn(self.st, mask, p, pos, rpy, cfg.spawn_throttle)
        self.step_count[mask] = 0  # hover.py:98
        self.terminated[mask] = False
        self.truncated[mask] = False
        self.action[mask] = 0.0  # hover.py:101
        self.info["out_of_bounds"][mask] = False
        self.info["on_floor"][mask] = False
        self.prev_centre[mask] = 0.0  # hover.py:105-107
        self.prev_area[mask] = 0.0
        self.prev_ratio[mask] = 0.0
        self.ep_return[mask] = 0.0
        # 10 idle Aviary.step() with the default (zero) setpoint, hover.py:109-110
        sub_state = self.st.select(mask)
        key = NoiseSource(self.noise.seed, self.env_ids[mask], self.noise.enabled)
        ctr = self.rng_ctr[mask]
        sp = mode_preset_setpoint(sub_state, p)  # zeros in mode 0
        sub = 0
        for _ in range(cfg.reset_idle_steps):
            sub = aviary_step(sub_state, sp, p, key.normals, sub, STREAM_RESET, ctr)
        for k, v in self.st.__dict__.items():
            v[mask] = sub_state.__dict__[k]
        self.prev_euler[mask] = self.st.s_euler[mask]  # hover.py:112

    def reset(self, mask: np.ndarray | None = None) -> np.ndarray:
        mask = np.ones(self.n, bool) if mask is None else np.asarray(mask, bool)
        self._reset_envs(mask)
        # compute_state() (hover.py:113) touches prev_* of every env, so only
        # splice the masked rows back.
        keep = (self.prev_centre.copy(), self.prev_area.copy(), self.prev_ratio.copy())
        obs = self._compute_state()
        if not mask.all():
            nm = ~mask
            self.prev_centre[nm], self.prev_area[nm], self.prev_ratio[nm] = keep[0][nm], keep[1][nm], keep[2][nm]
        self._last_obs = obs
        return obs

    # -------------------------------------------------------------------- step
    def step(self, actions: np.ndarray):
        """hover.py:334-358 for every env, then SB3 VecEnv auto-reset."""
        cfg, p = self.cfg, self.p
        a = np.asarray(actions, float).reshape(self.n, 4)
        self.action = a.copy()  # hover.py:335
        sp = np.stack(
            [a[:, 0] * cfg.action_scale[0], a[:, 1] * cfg.action_scale[1], a[:, 2] * cfg.action_scale[2],
             a[:, 3] * cfg.thrust_scale + cfg.thrust_bias],
            axis=1,
        )  # hover.py:337-341
        if p.flight_mode == -1:  # four motor pwm commands: all of them through the throttle mapping
            sp = a * cfg.thrust_scale + cfg.thrust_bias
        reward = np.full(self.n, -0.1)  # hover.py:343
        live = ~(self.terminated | self.truncated)  # hover.py:347-348
        if live.any():
            sub_state = self.st.select(live)
            key = NoiseSource(self.noise.seed, self.env_ids[live], self.noise.enabled)
            ctr = self.rng_ctr[live]
            sub = 0
            for _ in range(cfg.env_step_ratio):  # hover.py:346-349
                sub = aviary_step(sub_state, sp[live], p, key.normals, sub, STREAM_STEP, ctr)
            for k, v in self.st.__dict__.items():
                v[live] = sub_state.__dict__[k]
        self.rng_ctr += np.uint64(1)
        obs = self._compute_state()  # hover.py:351
        # ---- compute_term_trunc_reward, hover.py:274-332
        k = self.step_count
        self.truncated |= k > cfg.max_steps  # hover.py:275-276
        oob = np.linalg.norm(self.st.s_pos, axis=1) > cfg.flight_dome_size  # hover.py:278
        reward = np.where(oob, -100.0, reward)
        self.info["out_of_bounds"] |= oob
        self.terminated |= oob
        if not cfg.render:
            floor = (k > cfg.floor_grace_steps) & (self.st.s_pos[:, 2] < cfg.floor_threshold)  # hover.py:283-290
            reward = np.where(floor, -100.0, reward)
            self.info["on_floor"] |= floor
            self.terminated |= floor
        visible = obs[:, 13] > 0.5  # hover.py:296
        centre_d = np.sqrt(obs[:, 7] ** 2 + obs[:, 8] ** 2)  # hover.py:305
        area_d = np.abs(obs[:, 11] - cfg.target_area)  # hover.py:309
        ratio_d = np.abs(obs[:, 14] - cfg.target_ratio)  # hover.py:313
        target_reward = np.where(visible, (-centre_d) + (-area_d) + (-ratio_d), -2.0)  # hover.py:317,320
        yaw_rate = np.abs(self.st.s_wb[:, 2])  # hover.py:322
        reward = reward - 0.01 * yaw_rate**2  # hover.py:323-324
        ang_d = np.sqrt(self.st.s_euler[:, 0] ** 2 + self.st.s_euler[:, 1] ** 2)  # hover.py:326
        reward = reward + (target_reward - ang_d)  # hover.py:327
        smooth = np.sqrt(((self.action - self.prev_action) ** 2).sum(axis=1))  # hover.py:329-330
        reward = reward - smooth * 0.2  # hover.py:331
        reward = reward + 1.0  # hover.py:332
        # ---- back in step(), hover.py:354-357
        self.prev_euler = self.st.s_euler.copy()
        self.step_count = self.step_count + 1
        self.prev_action = self.action.copy()
        terminated, truncated = self.terminated.copy(), self.truncated.copy()
        info = {
            "out_of_bounds": self.info["out_of_bounds"].copy(),
            "on_floor": self.info["on_floor"].copy(),
            "terminal_obs": obs.copy(),
        }
        self.ep_return += reward
        done = terminated | truncated
        if self.auto_reset and done.any():
            self.sum_ret += float(self.ep_return[done].sum())
            self.sum_len += int(self.step_count[done].sum())
            self.n_done += int(done.sum())
            keep = (self.prev_centre.copy(), self.prev_area.copy(), self.prev_ratio.copy())
            self._reset_envs(done)
            obs_r = self._compute_state()
            nd = ~done
            self.prev_centre[nd], self.prev_area[nd], self.prev_ratio[nd] = keep[0][nd], keep[1][nd], keep[2][nd]
            obs = np.where(done[:, None], obs_r, obs)
        self._last_obs = obs
        return obs, reward, terminated, truncated, info
