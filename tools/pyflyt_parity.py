#!/usr/bin/env python
"""Settle the PyFlyt / pybullet unknowns where `pyflyt==0.21.0` can be installed.

Cannot run in the build sandbox (PyFlyt, pybullet and the network are absent) -- it is kept so that anyone with

    pip install pyflyt==0.21.0 pybullet==3.2.7

can run, from the reference's `simulation/` directory (for ./drone_models/cf2x):

    python /path/to/repo/tools/pyflyt_parity.py --out pyflyt_traj.npz
    python /path/to/repo/tools/pyflyt_parity.py --compare pyflyt_traj.npz

(add `--mode M` to both calls for the other flight modes, -1..7: the outer PID loops of cf2x.yaml:21-54, which
hover.py never reaches but `QxConfig.flight_mode` exposes.)

The first call steps the REAL PyFlyt Aviary (mode 0, noise_ratio patched to 0, start z = 1 m, free flight) through
fixed setpoint sequences and dumps Aviary.state after every Aviary.step; the second steps the restated oracle through
the same sequences and reports, per switch setting, the worst relative deviation over 500 control steps.  The switch
combination with the smallest deviation is the one `QxConfig` / `QuadXParams` should default to.
"""
from __future__ import annotations

import argparse
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# per flight mode: scale of the first three setpoint channels, (centre, half-range) of the fourth
MODE_SETPOINTS = {
    -1: ((0.0, 0.0, 0.0), (0.4952, 0.01)),   # motor pwm: all four near hover (channels 0-2 are set from channel 3 below)
    0: ((1.5, 1.5, 2.0), (0.4952, 0.03)),    # vp, vq, vr, thrust
    1: ((0.3, 0.3, 1.0), (0.0, 0.3)),        # p, q, r, vz
    2: ((1.0, 1.0, 1.0), (1.2, 0.3)),        # vp, vq, vr, z
    3: ((0.3, 0.3, 1.0), (1.2, 0.3)),        # p, q, r, z
    4: ((0.8, 0.8, 1.0), (1.2, 0.3)),        # u, v, vr, z
    5: ((0.8, 0.8, 1.0), (0.0, 0.3)),        # u, v, vr, vz
    6: ((0.8, 0.8, 1.0), (0.0, 0.3)),        # vx, vy, vr, vz
    7: ((1.0, 1.0, 1.5), (1.2, 0.3)),        # x, y, r, z
}


def setpoints(n_steps: int, seed: int = 0, mode: int = 0) -> np.ndarray:
    """Piecewise-constant commands in the units of the flight mode, gentle enough to stay airborne."""
    rng = np.random.default_rng(seed)
    scale, (centre, half) = MODE_SETPOINTS[mode]
    sp = np.zeros((n_steps, 4))
    hold = 20 if mode <= 0 else 120  # the outer loops need a second or so to act on a command
    for k in range(0, n_steps, hold):
        sp[k:k + hold, :3] = rng.uniform(-1.0, 1.0, 3) * np.array(scale)
        sp[k:k + hold, 3] = centre + rng.uniform(-half, half)
        if mode == -1:
            sp[k:k + hold, :3] = sp[k, 3] + rng.uniform(-0.002, 0.002, 3)
    return sp


def run_pyflyt(out: str, n_steps: int, mode: int = 0) -> None:
    from PyFlyt.core import Aviary  # noqa: the real one

    env = Aviary(start_pos=np.array([[0.0, 0.0, 1.0]]), start_orn=np.zeros((1, 3)), render=False, drone_type="quadx", physics_hz=240.0,
                 drone_options={"use_camera": False, "model_dir": "./drone_models", "drone_model": "cf2x"})
    env.reset()
    env.set_mode(mode)  # after reset(): QuadX.reset() puts every drone back into mode 0
    env.drones[0].motors.noise_ratio *= 0.0
    sp = setpoints(n_steps, mode=mode)
    states = []
    for k in range(n_steps):
        env.set_setpoint(0, sp[k])
        env.step()
        states.append(env.state(0).copy())
    np.savez_compressed(out, setpoints=sp, states=np.asarray(states), mode=mode, motor_map=np.asarray(env.drones[0].motor_map),
                        control_hz=getattr(env.drones[0], "control_period", 0.0))
    print("wrote", out)


def compare(path: str) -> None:
    from oracle.quadx_model import QuadXParams, QuadXState, aviary_step, spawn

    ref = np.load(path)
    sp, states = ref["setpoints"], ref["states"]
    mode = int(ref["mode"]) if "mode" in ref else 0
    best = None
    for stale, gyro, chz in itertools.product((True, False), (True, False), (120.0, 240.0)):
        p = QuadXParams(state_stale=stale, gyro=gyro, control_hz=chz, noise_ratio=0.0, flight_mode=mode)
        st = QuadXState.zeros(1)
        spawn(st, np.ones(1, bool), p, [[0.0, 0.0, 1.0]], [[0.0, 0.0, 0.0]], 0.0)
        worst = 0.0
        for k in range(sp.shape[0]):
            aviary_step(st, sp[k][None], p, None, 0, 0, np.zeros(1, np.uint64))
            err = np.abs(st.aviary_state()[0] - states[k]) / np.maximum(1.0, np.abs(states[k]))
            worst = max(worst, float(err.max()))
        print(f"state_stale={stale!s:5} gyro={gyro!s:5} control_hz={chz:5.0f}: worst rel deviation {worst:.3e}")
        if best is None or worst < best[0]:
            best = (worst, stale, gyro, chz)
    print("best:", best)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--compare", default="")
    ap.add_argument("--steps", type=int, default=3000, help="Aviary.step() calls (6 per hover control step)")
    ap.add_argument("--mode", type=int, default=0, choices=sorted(MODE_SETPOINTS), help="PyFlyt flight mode (QuadX.set_mode)")
    a = ap.parse_args()
    if a.out:
        run_pyflyt(a.out, a.steps, a.mode)
    if a.compare:
        compare(a.compare)
