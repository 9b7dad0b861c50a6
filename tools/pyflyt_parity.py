#!/usr/bin/env python
"""Settle the PyFlyt / pybullet unknowns where `pyflyt==0.21.0` can be installed.

Cannot run in the build sandbox (PyFlyt, pybullet and the network are absent) -- it is kept so that anyone with

    pip install pyflyt==0.21.0 pybullet==3.2.7

can run, from the reference's `simulation/` directory (for ./drone_models/cf2x):

    python /path/to/repo/tools/pyflyt_parity.py --out pyflyt_traj.npz
    python /path/to/repo/tools/pyflyt_parity.py --compare pyflyt_traj.npz

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


def setpoints(n_steps: int, seed: int = 0) -> np.ndarray:
    """Piecewise-constant body-rate + thrust commands (PyFlyt mode 0), gentle enough to stay airborne."""
    rng = np.random.default_rng(seed)
    sp = np.zeros((n_steps, 4))
    for k in range(0, n_steps, 20):
        sp[k:k + 20, :3] = rng.uniform(-1.0, 1.0, 3) * np.array([1.5, 1.5, 2.0])
        sp[k:k + 20, 3] = 0.4952 + rng.uniform(-0.03, 0.03)
    return sp


def run_pyflyt(out: str, n_steps: int) -> None:
    from PyFlyt.core import Aviary  # noqa: the real one

    env = Aviary(start_pos=np.array([[0.0, 0.0, 1.0]]), start_orn=np.zeros((1, 3)), render=False, drone_type="quadx", physics_hz=240.0,
                 drone_options={"use_camera": False, "model_dir": "./drone_models", "drone_model": "cf2x"})
    env.set_mode(0)
    env.reset()
    env.drones[0].motors.noise_ratio *= 0.0
    sp = setpoints(n_steps)
    states = []
    for k in range(n_steps):
        env.set_setpoint(0, sp[k])
        env.step()
        states.append(env.state(0).copy())
    np.savez_compressed(out, setpoints=sp, states=np.asarray(states), motor_map=np.asarray(env.drones[0].motor_map),
                        control_hz=getattr(env.drones[0], "control_period", 0.0))
    print("wrote", out)


def compare(path: str) -> None:
    from oracle.quadx_model import QuadXParams, QuadXState, aviary_step, spawn

    ref = np.load(path)
    sp, states = ref["setpoints"], ref["states"]
    best = None
    for stale, gyro, chz in itertools.product((True, False), (True, False), (120.0, 240.0)):
        p = QuadXParams(state_stale=stale, gyro=gyro, control_hz=chz, noise_ratio=0.0)
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
    a = ap.parse_args()
    if a.out:
        run_pyflyt(a.out, a.steps)
    if a.compare:
        compare(a.compare)
