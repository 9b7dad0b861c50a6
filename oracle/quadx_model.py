"""Batched float64 restatement of PyFlyt 0.21.0 ``QuadX`` (mode 0, the only mode
hover.py:92 reaches, plus the rest of its PID cascade: flight modes -1..7) on top of a
free-flight restatement of pybullet 3.2.7's multibody integrator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** for this
file: PyFlyt / pybullet are third-party dependencies pinned in
``/root/reference/simulation/uv.lock:807-829`` and absent from the reference
tree; what is restated here is their published algorithm as called from
``/root/reference/simulation/hover.py:77-93,110,344,349`` and parameterised by
``simulation/drone_models/cf2x/cf2x.yaml:1-53`` and ``cf2x.urdf:10-68``.

All arrays carry a leading env axis ``N``.  Quaternions are (x, y, z, w) like
pybullet.  Angular velocity ``omega`` is stored in the world frame (Bullet's
``btMultiBody::getBaseOmega``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------


@dataclass
class QuadXParams:
    """cf2x drone + every switch the restatement is unsure about."""

    # cf2x.yaml:2-6 motor_params
    total_thrust: float = 4.0
    thrust_coef: float = 3.16e-10
    torque_coef: float = 7.94e-12
    noise_ratio: float = 0.02
    tau: float = 0.01
    # cf2x.yaml:9-11 drag_params
    drag_coef_xyz: float = 1.5
    drag_area_xyz: float = 3.0e-4
    drag_coef_pqr: float = 1.0e-4
    air_density: float = 1.225  # PyFlyt BoringBodies constant [RECALL]
    # cf2x.yaml:14-19 control_params.ang_vel
    rate_kp: tuple = (2.0e-2, 2.0e-2, 4.0e-2)
    rate_ki: tuple = (2.5e-7, 2.5e-7, 1.35e-4)
    rate_kd: tuple = (5.0e-5, 5.0e-5, 0.0)
    rate_lim: tuple = (1.0, 1.0, 1.0)
    # cf2x.yaml:21-53: the rest of the cascade (unused by hover.py, which hard-codes set_mode(0) at :92)
    flight_mode: int = 0  # PyFlyt QuadX.set_mode: -1 pwm | 0 vp,vq,vr,T | 1 p,q,r,vz | 2 vp,vq,vr,z | 3 p,q,r,z
    #                        | 4 u,v,vr,z | 5 u,v,vr,vz | 6 vx,vy,vr,vz | 7 x,y,r,z
    att_kp: tuple = (1.0, 1.0, 1.0)  # ang_pos, cf2x.yaml:21-26
    att_ki: tuple = (0.0, 0.0, 0.0)
    att_kd: tuple = (0.0, 0.0, 0.0)
    att_lim: tuple = (2.0, 2.0, 2.0)
    vel_kp: tuple = (0.4, 0.4)  # lin_vel, cf2x.yaml:28-33
    vel_ki: tuple = (0.15, 0.15)
    vel_kd: tuple = (0.25, 0.25)
    vel_lim: tuple = (0.3, 0.3)
    pos_kp: tuple = (0.5, 0.5)  # lin_pos, cf2x.yaml:35-40
    pos_ki: tuple = (0.0, 0.0)
    pos_kd: tuple = (0.0, 0.0)
    pos_lim: tuple = (1.0, 1.0)
    zpos_pid: tuple = (0.8, 0.0, 0.0, 1.5)  # (kp, ki, kd, lim), cf2x.yaml:42-47
    zvel_pid: tuple = (1.5, 0.3, 0.05, 1.0)  # cf2x.yaml:49-54
    # cf2x.urdf:10,12 (U6: inertia taken from the file)
    mass: float = 0.1
    inertia: tuple = (3.0e-5, 3.0e-5, 5.0e-5)
    # cf2x.urdf:35,46,57,68 prop link positions (x, y); z = 0
    motor_xy: tuple = ((0.028, -0.028), (-0.028, 0.028), (0.028, 0.028), (-0.028, -0.028))
    # sign of the reaction torque of motors 0..3 about body z (U4)
    torque_sign: tuple = (-1.0, -1.0, 1.0, 1.0)
    # cmd (roll, pitch, yaw, thrust) -> pwm of motors 0..3 (U4); rows = motors.
    # Derived from the link positions above: tau = r x F = (y F, -x F, 0).
    motor_map: tuple = (
        (-1.0, -1.0, -1.0, 1.0),
        (+1.0, +1.0, -1.0, 1.0),
        (+1.0, -1.0, +1.0, 1.0),
        (-1.0, +1.0, +1.0, 1.0),
    )
    pwm_idle: float = 0.05
    # scheduler: hover.py:23 physics_hz; PyFlyt QuadX default control_hz (U1)
    physics_hz: float = 240.0
    control_hz: float = 120.0
    gravity: float = 9.81
    # U2: Aviary.step order is control -> physics -> update_state -> stepSimulation,
    # so state(i) / the controller input is one sub-step stale.
    state_stale: bool = True
    # U5: noise is applied after the lag, multiplicative on the throttle.
    # U7: drag = -sign(v) * 0.5 rho Cd A v^2 (body frame); rotational drag
    #     skipped while in contact.
    # U8: gyroscopic term on.
    gyro: bool = True
    # U9: btMultiBody m_maxCoordinateVelocity clamp on every base velocity dof.
    max_coord_vel: float = 100.0
    # U10: declared floor stand-in (no contact solver): sticky plane at the
    # rest height of the 0.02 m tall collision box (cf2x.urdf:29).
    floor_z: float = 0.01
    # U11 / hover.py:85-87 camera; camera_angle_degrees=-25 is read as 25 deg UP.
    cam_tilt_up_deg: float = 25.0
    cam_fov_deg: float = 90.0
    cam_res: int = 128
    cam_near: float = 0.1
    cam_offset: tuple = (0.0, 0.0, 0.0)

    # derived ---------------------------------------------------------------
    @property
    def h(self) -> float:
        return 1.0 / self.physics_hz

    @property
    def ctrl_period(self) -> float:
        return 1.0 / self.control_hz

    @property
    def substeps_per_aviary_step(self) -> int:
        return int(self.physics_hz / self.control_hz)

    @property
    def max_rpm(self) -> float:
        return float(np.sqrt(self.total_thrust / (4.0 * self.thrust_coef)))

    @property
    def drag_const(self) -> float:
        return 0.5 * self.air_density * self.drag_coef_xyz * self.drag_area_xyz


# ----------------------------------------------------------------------------
# Philox4x32-R counter RNG (identical integer stream in numpy, C and CUDA).  The env's noise uses R = 7 rounds
# (Random123: the smallest round count that passes BigCrush; csrc/qx_model.cuh kEnvPhiloxRounds).
# ----------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)

STREAM_STEP = 0  # motor noise during an agent step
STREAM_RESET = 1  # motor noise during the 10 idle Aviary.step() of a reset
STREAM_SPAWN = 2  # reset pose noise


ENV_PHILOX_ROUNDS = 7


def philox4x32(ctr: np.ndarray, key: np.ndarray, rounds: int = ENV_PHILOX_ROUNDS) -> np.ndarray:
    """ctr uint32[N,4], key uint32[N,2] -> uint32[N,4]."""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0 = key[:, 0].astype(np.uint32).copy()
    k1 = key[:, 1].astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for r in range(rounds):
            if r > 0:
                k0 = (k0 + _W0).astype(np.uint32)
                k1 = (k1 + _W1).astype(np.uint32)
            p0 = _M0 * c[0]
            p1 = _M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
    return np.stack([x.astype(np.uint32) for x in c], axis=1)


def rng_key(seed: int, env_ids: np.ndarray) -> np.ndarray:
    """Per-env Philox key: (seed_lo ^ env_id_lo, seed_hi ^ env_id_hi)."""
    env_ids = env_ids.astype(np.uint64)
    k0 = (np.uint64(seed & 0xFFFFFFFF) ^ (env_ids & _MASK)).astype(np.uint32)
    k1 = (np.uint64((seed >> 32) & 0xFFFFFFFF) ^ (env_ids >> np.uint64(32))).astype(np.uint32)
    return np.stack([k0, k1], axis=1)


def u01(x: np.ndarray) -> np.ndarray:
    """uint32 -> (0,1): ((x >> 9) + 0.5) * 2^-23 -- exact in fp32 and fp64 (on
    the GPU: bit-cast (x >> 9) | 0x3f800000 and subtract 1 - 2^-24, no I2F).
    Used for the reset pose noise."""
    return ((x >> np.uint32(9)).astype(np.float64) + 0.5) * (1.0 / 8388608.0)


def u01_16(x: np.ndarray) -> np.ndarray:
    """16-bit field (as uint32) -> (0,1): (x + 0.5) * 2^-16."""
    return (x.astype(np.float64) + 0.5) * (1.0 / 65536.0)


def normal4(w: np.ndarray) -> np.ndarray:
    """Two 32-bit words [N,2] -> four N(0,1) [N,4]: each word feeds one
    Box-Muller pair, low 16 bits -> radius, high 16 bits -> angle; outputs are
    (r_a cos, r_a sin, r_b cos, r_b sin) for motors 0..3, with the angle 2 pi u - pi in (-pi, pi) (where the GPU's
    MUFU sine / cosine are most accurate).  One Philox call (4 words) therefore serves two physics sub-steps."""
    lo = w & np.uint32(0xFFFF)
    hi = w >> np.uint32(16)
    r = np.sqrt(-2.0 * np.log(u01_16(lo)))
    t = 2.0 * np.pi * u01_16(hi) - np.pi
    return np.stack([r[:, 0] * np.cos(t[:, 0]), r[:, 0] * np.sin(t[:, 0]), r[:, 1] * np.cos(t[:, 1]), r[:, 1] * np.sin(t[:, 1])], axis=1)


# ----------------------------------------------------------------------------
# rotation helpers (pybullet conventions)
# ----------------------------------------------------------------------------


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    """(x,y,z,w)[N,4] -> body->world rotation [N,3,3]."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def quat_to_euler(q: np.ndarray) -> np.ndarray:
    """pybullet.getEulerFromQuaternion (btQuaternion::getEulerZYX) -> (roll, pitch, yaw)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    sarg = -2.0 * (x * z - w * y)
    roll = np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
    yaw = np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    lo = sarg <= -0.99999
    hi = sarg >= 0.99999
    if lo.any() or hi.any():
        roll = np.where(lo | hi, 0.0, roll)
        pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
        yaw = np.where(lo, 2 * np.arctan2(x, -y), np.where(hi, 2 * np.arctan2(-x, y), yaw))
    return np.stack([roll, pitch, yaw], axis=1)


def euler_to_quat(e: np.ndarray) -> np.ndarray:
    """pybullet.getQuaternionFromEuler (hover.py:233): (roll,pitch,yaw)[N,3] -> (x,y,z,w)."""
    hr, hp, hy = 0.5 * e[:, 0], 0.5 * e[:, 1], 0.5 * e[:, 2]
    cr, sr = np.cos(hr), np.sin(hr)
    cp, sp = np.cos(hp), np.sin(hp)
    cy, sy = np.cos(hy), np.sin(hy)
    return np.stack(
        [
            sr * cp * cy - cr * sp * sy,
            cr * sp * cy + sr * cp * sy,
            cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy,
        ],
        axis=1,
    )


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return np.stack(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        axis=1,
    )


# ----------------------------------------------------------------------------
# batched drone state + one physics sub-step
# ----------------------------------------------------------------------------


# columns of QuadXState.cpid: (integral, previous error) of ang_pos[3], lin_vel[2], lin_pos[2], z_vel, z_pos
CP_ATT_I, CP_ATT_E, CP_VEL_I, CP_VEL_E, CP_POS_I, CP_POS_E, CP_ZV_I, CP_ZV_E, CP_ZP_I, CP_ZP_E = 0, 3, 6, 8, 10, 12, 14, 15, 16, 17
CP_WORDS = 18


@dataclass
class QuadXState:
    """True rigid-body state + controller/motor state + the (possibly stale)
    snapshot PyFlyt exposes as ``Aviary.state(i)``."""

    pos: np.ndarray
    quat: np.ndarray
    vel: np.ndarray
    omega: np.ndarray  # world frame
    thr: np.ndarray  # motor throttle [N,4]
    pid_i: np.ndarray
    pid_e: np.ndarray
    contact: np.ndarray  # bool[N]
    # snapshot (QuadX.update_state): rows of Aviary.state(i)
    s_wb: np.ndarray
    s_euler: np.ndarray
    s_vb: np.ndarray
    s_pos: np.ndarray
    # integral / previous error of the outer loops (flight modes 1..7), columns CP_*
    cpid: np.ndarray = None

    def __post_init__(self):
        if self.cpid is None:
            self.cpid = np.zeros((self.pos.shape[0], CP_WORDS))

    @classmethod
    def zeros(cls, n: int) -> "QuadXState":
        z3 = lambda: np.zeros((n, 3))  # noqa: E731
        q = np.zeros((n, 4))
        q[:, 3] = 1.0
        return cls(z3(), q, z3(), z3(), np.zeros((n, 4)), z3(), z3(), np.zeros(n, bool), z3(), z3(), z3(), z3())

    def copy(self) -> "QuadXState":
        return QuadXState(**{k: v.copy() for k, v in self.__dict__.items()})

    def select(self, idx) -> "QuadXState":
        return QuadXState(**{k: v[idx].copy() for k, v in self.__dict__.items()})

    def assign(self, mask: np.ndarray, other: "QuadXState") -> None:
        for k, v in self.__dict__.items():
            v[mask] = other.__dict__[k][mask]

    def aviary_state(self) -> np.ndarray:
        """[N,4,3] rows (ang_vel body, euler, lin_vel body, pos) as read at
        hover.py:226,278,284,322,326."""
        return np.stack([self.s_wb, self.s_euler, self.s_vb, self.s_pos], axis=1)


def take_snapshot(st: QuadXState) -> None:
    """PyFlyt QuadX.update_state [RECALL]: body-frame velocities, Euler, position."""
    R = quat_to_mat(st.quat)
    st.s_wb = np.einsum("nji,nj->ni", R, st.omega)
    st.s_vb = np.einsum("nji,nj->ni", R, st.vel)
    st.s_euler = quat_to_euler(st.quat)
    st.s_pos = st.pos.copy()


def spawn(st: QuadXState, mask: np.ndarray, p: QuadXParams, pos, rpy, throttle: float = 0.0) -> None:
    """(Re)create the drone: Aviary(start_pos, start_orn) + Aviary.reset() (hover.py:77-93)."""
    n = int(mask.sum())
    if n == 0:
        return
    pos = np.broadcast_to(np.asarray(pos, float), (st.pos.shape[0], 3))[mask].copy()
    rpy = np.broadcast_to(np.asarray(rpy, float), (st.pos.shape[0], 3))[mask]
    pos[:, 2] = np.maximum(pos[:, 2], p.floor_z)
    st.pos[mask] = pos
    st.quat[mask] = euler_to_quat(rpy)
    st.vel[mask] = 0.0
    st.omega[mask] = 0.0
    st.thr[mask] = throttle
    st.pid_i[mask] = 0.0
    st.pid_e[mask] = 0.0
    st.cpid[mask] = 0.0
    st.contact[mask] = pos[:, 2] <= p.floor_z
    sub = st.select(mask)
    take_snapshot(sub)
    st.s_wb[mask], st.s_euler[mask], st.s_vb[mask], st.s_pos[mask] = sub.s_wb, sub.s_euler, sub.s_vb, sub.s_pos


def _pid(st: QuadXState, ci: int, ce: int, w: int, gains, T: float, state: np.ndarray, setpoint: np.ndarray) -> np.ndarray:
    """PyFlyt PID.step [RECALL]: clipped integral, derivative on the error, clipped sum; memory in st.cpid[:, ci:ci+w]."""
    kp, ki, kd, lim = (np.asarray(g, float) for g in gains)
    err = setpoint - state
    integ = np.clip(st.cpid[:, ci:ci + w] + ki * err * T, -lim, lim)
    d = kd * (err - st.cpid[:, ce:ce + w]) / T
    st.cpid[:, ci:ci + w] = integ
    st.cpid[:, ce:ce + w] = err
    return np.clip(kp * err + integ + d, -lim, lim)


def mode_preset_setpoint(st: QuadXState, p: QuadXParams) -> np.ndarray:
    """Setpoint PyFlyt's QuadX.set_mode leaves behind [RECALL] -- what the drone tracks during the idle
    Aviary.step()s of a reset (hover.py:109-110): hold the current height (modes 2, 3, 4) / pose (mode 7), else zeros."""
    sp = np.zeros((st.pos.shape[0], 4))
    if p.flight_mode in (2, 3, 4):
        sp[:, 3] = st.s_pos[:, 2]
    elif p.flight_mode == 7:
        sp[:, 0:2] = st.s_pos[:, 0:2]
        sp[:, 2] = st.s_euler[:, 2]
        sp[:, 3] = st.s_pos[:, 2]
    return sp


def control_update(st: QuadXState, setpoint: np.ndarray, p: QuadXParams) -> np.ndarray:
    """PyFlyt QuadX.update_control [RECALL]: outer loops of the flight mode -> rate PID -> mix -> saturation.
    Mode 0 (hover.py:92): setpoint [N,4] = (p, q, r rad/s, thrust 0..1).  Returns pwm [N,4]."""
    mode = p.flight_mode
    if mode == -1:  # direct motor commands, no mixing / saturation
        return setpoint.copy()
    kp, ki, kd, lim = (np.asarray(v) for v in (p.rate_kp, p.rate_ki, p.rate_kd, p.rate_lim))
    T = p.ctrl_period
    a_out = setpoint[:, :3].copy()
    z_out = setpoint[:, 3].copy()
    att2 = tuple(g[:2] for g in (p.att_kp, p.att_ki, p.att_kd, p.att_lim))
    if mode in (1, 3):  # angles -> rates
        a_out = _pid(st, CP_ATT_I, CP_ATT_E, 3, (p.att_kp, p.att_ki, p.att_kd, p.att_lim), T, st.s_euler, a_out)
    elif mode in (4, 5, 6, 7):
        if mode == 7:  # ground-frame position -> ground-frame velocity
            a_out[:, :2] = _pid(st, CP_POS_I, CP_POS_E, 2, (p.pos_kp, p.pos_ki, p.pos_kd, p.pos_lim), T, st.s_pos[:, :2], a_out[:, :2])
        if mode in (6, 7):  # ground frame -> local (yaw-rotated) frame
            c, s_ = np.cos(st.s_euler[:, 2]), np.sin(st.s_euler[:, 2])
            a_out[:, 0], a_out[:, 1] = c * a_out[:, 0] + s_ * a_out[:, 1], -s_ * a_out[:, 0] + c * a_out[:, 1]
        # local velocity -> tilt angles; +x needs +pitch, +y needs -roll
        ang = _pid(st, CP_VEL_I, CP_VEL_E, 2, (p.vel_kp, p.vel_ki, p.vel_kd, p.vel_lim), T, st.s_vb[:, :2], a_out[:, :2])
        a_out[:, 0], a_out[:, 1] = -ang[:, 1], ang[:, 0]
        if mode == 7:  # all three are angles
            a_out = _pid(st, CP_ATT_I, CP_ATT_E, 3, (p.att_kp, p.att_ki, p.att_kd, p.att_lim), T, st.s_euler, a_out)
        else:  # roll / pitch angles, yaw stays a rate
            a_out[:, :2] = _pid(st, CP_ATT_I, CP_ATT_E, 2, att2, T, st.s_euler[:, :2], a_out[:, :2])
    # height: z -> vz -> thrust
    if mode in (2, 3, 4, 7):
        z_out = _pid(st, CP_ZP_I, CP_ZP_E, 1, tuple((g,) for g in p.zpos_pid), T, st.s_pos[:, 2:3], z_out[:, None])[:, 0]
    if mode != 0:
        z_out = _pid(st, CP_ZV_I, CP_ZV_E, 1, tuple((g,) for g in p.zvel_pid), T, st.s_vb[:, 2:3], z_out[:, None])[:, 0]
    z_out = np.clip(z_out, 0.0, 1.0)
    e = a_out - st.s_wb
    st.pid_i = np.clip(st.pid_i + ki * e * T, -lim, lim)
    d = kd * (e - st.pid_e) / T
    out = np.clip(kp * e + st.pid_i + d, -lim, lim)
    st.pid_e = e
    cmd = np.concatenate([out, z_out[:, None]], axis=1)
    pwm = cmd @ np.asarray(p.motor_map).T
    high = pwm.max(axis=1, keepdims=True)
    pwm = np.where(high > 1.0, pwm / np.where(high > 1.0, high, 1.0), pwm)
    low = pwm.min(axis=1, keepdims=True)
    fix = pwm + (1.0 - pwm) / (1.0 - np.where(low < p.pwm_idle, low, 0.0)) * (p.pwm_idle - low)
    return np.where(low < p.pwm_idle, fix, pwm)


def physics_substep(st: QuadXState, pwm: np.ndarray, normals: np.ndarray | None, p: QuadXParams) -> None:
    """QuadX.update_physics -> update_state -> pybullet.stepSimulation for one
    1/physics_hz sub-step [RECALL], free flight + the declared floor stand-in."""
    h = p.h
    # --- Motors.physics_update: first-order lag, multiplicative noise, thrust/torque
    st.thr = st.thr + (h / p.tau) * (pwm - st.thr)
    if normals is not None and p.noise_ratio != 0.0:
        st.thr = st.thr + normals * st.thr * p.noise_ratio
    rpm = st.thr * p.max_rpm
    rr = np.abs(rpm) * rpm
    thrust = p.thrust_coef * rr
    mxy = np.asarray(p.motor_xy)
    tq = p.torque_coef * rr * np.asarray(p.torque_sign)
    # --- drag from the (stale) snapshot
    c = p.drag_const
    f_b = -np.sign(st.s_vb) * c * st.s_vb**2
    f_b[:, 2] += thrust.sum(axis=1)
    tau_b = np.stack([(mxy[:, 1] * thrust).sum(axis=1), -(mxy[:, 0] * thrust).sum(axis=1), tq.sum(axis=1)], axis=1)
    drag_pqr = -p.drag_coef_pqr * st.s_wb**2 * np.sign(st.s_wb)
    tau_b = tau_b + np.where(st.contact[:, None], 0.0, drag_pqr)
    # --- update_state happens before stepSimulation (U2)
    if p.state_stale:
        take_snapshot(st)
    # --- stepSimulation: semi-implicit Euler on the floating base
    R = quat_to_mat(st.quat)
    I = np.asarray(p.inertia)
    wb = np.einsum("nji,nj->ni", R, st.omega)
    gyro = np.cross(wb, I * wb) if p.gyro else 0.0
    wdot_b = (tau_b - gyro) / I
    acc = np.einsum("nij,nj->ni", R, f_b) / p.mass
    acc[:, 2] -= p.gravity
    st.omega = np.clip(st.omega + h * np.einsum("nij,nj->ni", R, wdot_b), -p.max_coord_vel, p.max_coord_vel)
    st.vel = np.clip(st.vel + h * acc, -p.max_coord_vel, p.max_coord_vel)
    st.pos = st.pos + h * st.vel
    # exponential-map quaternion update (btMultiBody::stepPositionsMultiDof)
    ang = np.linalg.norm(st.omega, axis=1)
    half = 0.5 * ang * h
    small = ang < 1e-3
    k = np.where(small, 0.5 * h - (h**3) * 0.020833333333 * ang * ang, np.sin(half) / np.where(small, 1.0, ang))
    dq = np.concatenate([st.omega * k[:, None], np.cos(half)[:, None]], axis=1)
    q = quat_mul(dq, st.quat)
    st.quat = q / np.linalg.norm(q, axis=1, keepdims=True)
    # --- floor stand-in (U10)
    below = st.pos[:, 2] < p.floor_z
    st.pos[:, 2] = np.where(below, p.floor_z, st.pos[:, 2])
    st.vel[:, 2] = np.where(below, np.maximum(st.vel[:, 2], 0.0), st.vel[:, 2])
    st.vel[:, :2] = np.where(below[:, None], 0.0, st.vel[:, :2])
    st.omega[:, :2] = np.where(below[:, None], 0.0, st.omega[:, :2])
    st.contact = below
    if not p.state_stale:
        take_snapshot(st)


@dataclass
class NoiseSource:
    """Counter-based motor noise: Philox key per env, counter
    (sub-step index, stream, env step counter lo, hi)."""

    seed: int
    env_ids: np.ndarray
    enabled: bool = True
    _key: np.ndarray = field(init=False)

    def __post_init__(self):
        self._key = rng_key(self.seed, np.asarray(self.env_ids))

    def bits(self, sub: int, stream: int, step_ctr: np.ndarray, mask=None) -> np.ndarray:
        n = self._key.shape[0]
        ctr = np.empty((n, 4), np.uint32)
        ctr[:, 0] = np.uint32(sub)
        ctr[:, 1] = np.uint32(stream)
        sc = step_ctr.astype(np.uint64)
        ctr[:, 2] = (sc & _MASK).astype(np.uint32)
        ctr[:, 3] = (sc >> np.uint64(32)).astype(np.uint32)
        return philox4x32(ctr, self._key)

    def normals(self, sub: int, stream: int, step_ctr: np.ndarray):
        """Motor noise of sub-step ``sub``: Philox counter word 0 is sub >> 1,
        the sub-step parity picks the word pair."""
        if not self.enabled:
            return None
        b = self.bits(sub >> 1, stream, step_ctr)
        return normal4(b[:, 2 * (sub & 1): 2 * (sub & 1) + 2])


def aviary_step(st: QuadXState, setpoint: np.ndarray, p: QuadXParams, noise, sub0: int, stream: int, step_ctr) -> int:
    """One PyFlyt ``Aviary.step()`` [RECALL]: ``physics_hz/control_hz`` sub-steps,
    control on the first.  ``noise(sub, stream, ctr)`` -> normals[N,4] or None.
    Returns the next sub-step index."""
    pwm = None
    for j in range(p.substeps_per_aviary_step):
        if j == 0:
            pwm = control_update(st, setpoint, p)
        physics_substep(st, pwm, noise(sub0 + j, stream, step_ctr) if noise is not None else None, p)
    return sub0 + p.substeps_per_aviary_step
