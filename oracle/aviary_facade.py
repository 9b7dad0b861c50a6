"""A stand-in for ``PyFlyt.core.Aviary`` exposing exactly the calls the
reference makes (hover.py:77-93,110,112,131-155,225,241,278,284,322,326,344,349,
354,361,365), backed by oracle/quadx_model.py and the CPU rasteriser.

TEST INFRASTRUCTURE ONLY.  It lets the reference's *own* ``hover.py`` run
unmodified (imported with stub modules, see ``install_stubs``) on top of the
restated drone model, which is how the env layer of the oracle and of the
CUDA path is pinned: same Aviary underneath, reference code on top.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import vision
from .quadx_model import STREAM_RESET, STREAM_STEP, NoiseSource, QuadXParams, QuadXState, aviary_step, euler_to_quat, spawn

# The reference builds a fresh Aviary on every reset() and never seeds it
# (hover.py:73-91), so the harness injects the noise schedule from outside:
#   NOISE_CONTEXT = {"source": NoiseSource | None, "rng_ctr": int,
#                    "idle_steps": 10, "ratio": 6, "params": QuadXParams}
NOISE_CONTEXT: dict = {"source": None, "rng_ctr": 0, "idle_steps": 10, "ratio": 6, "params": None}


class _Drone:
    def __init__(self, aviary: "Aviary"):
        self._aviary = aviary

    @property
    def rgbaImg(self) -> np.ndarray:  # hover.py:241
        a = self._aviary
        return vision.render_rgba(a._st.pos[0], a._st.quat[0], a._p, a._box_corners)


class Aviary:
    GEOM_BOX = 3  # pybullet.GEOM_BOX

    def __init__(self, start_pos, start_orn, drone_type="quadx", drone_options=None, render=False, physics_hz=240, **_):
        assert drone_type == "quadx" and np.asarray(start_pos).shape == (1, 3)
        opts = drone_options or {}
        p = NOISE_CONTEXT["params"] or QuadXParams()
        assert opts.get("drone_model", "cf2x") == "cf2x"
        assert float(physics_hz) == p.physics_hz
        # hover.py:84-87 -> camera parameters (tilt sign: U11)
        if "camera_angle_degrees" in opts:
            assert abs(opts["camera_angle_degrees"]) == p.cam_tilt_up_deg
        assert opts.get("camera_FOV_degrees", 90) == p.cam_fov_deg
        assert tuple(opts.get("camera_resolution", (128, 128))) == (p.cam_res, p.cam_res)
        self._p = p
        self._start_pos = np.asarray(start_pos, float)
        self._start_orn = np.asarray(start_orn, float)
        self._st = QuadXState.zeros(1)
        self._setpoint = np.zeros((1, 4))
        self._mode = None
        self._steps = 0
        self._visuals = {}
        self._box_corners = None
        self.drones = [_Drone(self)]

    # -- calls made by hover.py ------------------------------------------------
    def set_mode(self, mode):  # hover.py:92
        assert mode == 0, "only PyFlyt mode 0 (body rates + thrust) is restated"
        self._mode = 0
        self._setpoint = np.zeros((1, 4))

    def reset(self):  # hover.py:93
        spawn(self._st, np.ones(1, bool), self._p, self._start_pos, self._start_orn, 0.0)
        self._steps = 0

    def createVisualShape(self, shapeType, halfExtents, rgbaColor, visualFramePosition):  # hover.py:131-136
        assert shapeType == self.GEOM_BOX and list(rgbaColor) == [1, 0, 0, 1]
        vid = len(self._visuals)
        self._visuals[vid] = (np.asarray(halfExtents, float), np.asarray(visualFramePosition, float))
        return vid

    def createMultiBody(self, baseMass, baseVisualShapeIndex, basePosition, baseOrientation):  # hover.py:150-155
        assert baseMass == 0
        he, off = self._visuals[baseVisualShapeIndex]
        x, y, z, w = baseOrientation
        n = np.sqrt(x * x + y * y + z * z + w * w)
        x, y, z, w = x / n, y / n, z / n, w / n
        R = np.array(
            [
                [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
            ]
        )
        signs = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
        self._box_corners = np.asarray(basePosition, float) + (off + signs * he) @ R.T
        return 1

    def set_setpoint(self, index, setpoint):  # hover.py:344
        assert index == 0
        self._setpoint = np.asarray(setpoint, float).reshape(1, 4).copy()

    def step(self):  # hover.py:110,349
        ctx = NOISE_CONTEXT
        src: NoiseSource | None = ctx["source"]
        idle, ratio = ctx["idle_steps"], ctx["ratio"]
        per = self._p.substeps_per_aviary_step
        if self._steps < idle:
            stream, sub0 = STREAM_RESET, self._steps * per
        else:
            stream, sub0 = STREAM_STEP, ((self._steps - idle) % ratio) * per
        ctr = np.array([ctx["rng_ctr"]], np.uint64)
        fn = src.normals if (src is not None and src.enabled) else None
        aviary_step(self._st, self._setpoint, self._p, fn, sub0, stream, ctr)
        self._steps += 1

    def state(self, index):  # hover.py:112,225,278,...
        assert index == 0
        return self._st.aviary_state()[0].copy()

    def disconnect(self):  # hover.py:75,365
        pass

    def render(self):  # hover.py:361
        return self.drones[0].rgbaImg


def _get_quaternion_from_euler(e):
    """pybullet.getQuaternionFromEuler, called on the bare module at hover.py:233."""
    return tuple(euler_to_quat(np.asarray(e, float).reshape(1, 3))[0])


def install_stubs() -> None:
    """Register stand-ins for the third-party imports of hover.py:1-8 that are
    not installed here (gymnasium, PyFlyt, pybullet, radio_controller).  cv2 and
    numpy are the real packages."""

    class _Env:
        def reset(self, seed=None, options=None):
            return None

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float64):
            self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
            self.shape = self.low.shape if shape is None else shape

    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = _Box
    gym.Env = _Env
    gym.spaces = spaces
    pyflyt = types.ModuleType("PyFlyt")
    core = types.ModuleType("PyFlyt.core")
    core.Aviary = Aviary
    pyflyt.core = core
    pb = types.ModuleType("pybullet")
    pb.getQuaternionFromEuler = _get_quaternion_from_euler
    rc = types.ModuleType("radio_controller")
    rc.RadioMasterJoystick = type("RadioMasterJoystick", (), {})
    for name, mod in {
        "gymnasium": gym,
        "gymnasium.spaces": spaces,
        "PyFlyt": pyflyt,
        "PyFlyt.core": core,
        "pybullet": pb,
        "radio_controller": rc,
    }.items():
        sys.modules.setdefault(name, mod)


def import_reference_hover(reference_root: str = "/root/reference"):
    """Import the reference's hover.py unmodified (only possible where
    /root/reference exists, i.e. in the build container)."""
    import importlib
    import os

    install_stubs()
    sim = os.path.join(reference_root, "simulation")
    if sim not in sys.path:
        sys.path.insert(0, sim)
    return importlib.import_module("hover")
