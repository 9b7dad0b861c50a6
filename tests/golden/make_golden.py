"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own
``simulation/hover.py`` (imported unmodified from /root/reference with stub
modules for the packages that are not installed) on top of the restated drone
model (oracle/aviary_facade.py).

Run from the repo root, in the build container only:

    python tests/golden/make_golden.py

What the vectors pin: everything hover.py itself computes -- action scaling
(hover.py:337-341), the 6:1 step ratio (:346-349), observation packing
(:224-272), ``detect_rectangle`` on the rasterised frame (:157-222), reward and
termination (:274-332), the reset protocol (:72-114) incl. the un-reset
``prev_action`` quirk.  What they do NOT pin: PyFlyt/pybullet arithmetic
(parity unpinned, see oracle/__init__.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import aviary_facade as af  # noqa: E402
from oracle.quadx_model import NoiseSource, QuadXParams  # noqa: E402
from oracle.test_policy import hover_actions  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_scenario(hover, name, seed, noise, n_steps, action_fn, env_id=0, second_episode=0, render=False, agent_hz=40):
    p = QuadXParams()
    src = NoiseSource(seed, np.array([env_id], np.uint64), enabled=noise)
    af.NOISE_CONTEXT.update(source=src, rng_ctr=0, params=p, idle_steps=10, ratio=int(240 / agent_hz))
    env = hover.QuadXHoverEnv(agent_hz=agent_hz, render=render)
    rec = {k: [] for k in ("actions", "obs", "reward", "terminated", "truncated", "state", "episode_start")}
    reset_obs = []

    def episode(n):
        obs, info = env.reset()
        reset_obs.append(np.asarray(obs, np.float64))
        for k in range(n):
            a = action_fn(k, env.aviary.state(0))
            o, r, te, tr, inf = env.step(a.copy())
            af.NOISE_CONTEXT["rng_ctr"] += 1
            rec["actions"].append(a)
            rec["obs"].append(np.asarray(o, np.float64))
            rec["reward"].append(float(r))
            rec["terminated"].append(bool(te))
            rec["truncated"].append(bool(tr))
            rec["state"].append(env.aviary.state(0))
            rec["episode_start"].append(k == 0)

    episode(n_steps)
    if second_episode:
        episode(second_episode)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["reset_obs"] = np.asarray(reset_obs)
    out["seed"] = np.int64(seed)
    out["noise"] = np.bool_(noise)
    out["env_id"] = np.int64(env_id)
    out["render"] = np.bool_(render)
    out["agent_hz"] = np.int64(agent_hz)
    path = os.path.join(OUT, f"hover_ref_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(rec['reward'])} steps, terminated at {np.argmax(out['terminated']) if out['terminated'].any() else None}, "
          f"truncated at {np.argmax(out['truncated']) if out['truncated'].any() else None}, return {out['reward'].sum():.3f} -> {path}")


def drone_params(ref_root="/root/reference") -> dict:
    """The numbers of simulation/drone_models/cf2x/cf2x.yaml and cf2x.urdf, parsed from the
    reference's own files, as a flat dict of lists."""
    import xml.etree.ElementTree as ET

    import yaml

    d = os.path.join(ref_root, "simulation", "drone_models", "cf2x")
    y = yaml.safe_load(open(os.path.join(d, "cf2x.yaml")))
    out = {f"motor.{k}": float(v) for k, v in y["motor_params"].items()}
    out.update({f"drag.{k}": float(v) for k, v in y["drag_params"].items()})
    for loop, g in y["control_params"].items():
        for k in ("kp", "ki", "kd", "lim"):
            out[f"{loop}.{k}"] = [float(x) for x in np.atleast_1d(g[k])]
    base = ET.parse(os.path.join(d, "cf2x.urdf")).getroot().find("link[@name='base_link']")
    inert = base.find("inertial")
    out["urdf.mass"] = float(inert.find("mass").get("value"))
    out["urdf.inertia_diag"] = [float(inert.find("inertia").get(k)) for k in ("ixx", "iyy", "izz")]
    out["urdf.collision_box"] = [float(x) for x in base.find("collision/geometry/box").get("size").split()]
    root = ET.parse(os.path.join(d, "cf2x.urdf")).getroot()
    out["urdf.prop_xyz"] = [[float(x) for x in root.find(f"link[@name='prop{i}_link']/inertial/origin").get("xyz").split()] for i in (1, 2, 3, 4)]
    return out


def main():
    import json

    with open(os.path.join(OUT, "cf2x_params.json"), "w") as f:
        json.dump(drone_params(), f, indent=1, sort_keys=True)
    hover = af.import_reference_hover("/root/reference")
    rng = np.random.default_rng(2024)
    tgt = np.array([[0.4, -0.3, 0.9]])

    def fly(k, state):
        return hover_actions(state[None], tgt, rng, 0.05)[0]

    def idle(k, state):
        return np.array([0.0, 0.0, 0.0, -1.0])

    def rocket(k, state):
        return np.array([0.02, -0.03, 0.01, 1.0])

    # (a) take-off from the floor, hover, run into the time limit (402 steps) and 6 steps past it
    run_scenario(hover, "fly_quiet", seed=11, noise=False, n_steps=408, action_fn=fly)
    # (b) same with 2 % motor noise, followed by a second episode (prev_action carries over, hover.py:31,357)
    run_scenario(hover, "fly_noisy", seed=1234, noise=True, n_steps=120, action_fn=fly, env_id=5, second_episode=40)
    # (c) never takes off: floor rule fires on the 32nd step (hover.py:283-290), then sticky flags
    run_scenario(hover, "floor", seed=7, noise=True, n_steps=40, action_fn=idle)
    # (d) full throttle: leaves the 3 m dome (hover.py:278-281)
    run_scenario(hover, "dome", seed=9, noise=False, n_steps=40, action_fn=rocket)
    # (e) render=True switches the floor rule off (hover.py:283 `and not self.render`): the same idle drone is never terminated
    run_scenario(hover, "render_idle", seed=7, noise=True, n_steps=40, action_fn=idle, render=True)
    # (f) agent_hz=60 (hover.py:14,24-25): env_step_ratio = int(240 / 60) = 4 Aviary.step per agent step, agent_dt = 1/60
    # (g) actions outside the Box(-1, 1): hover.py neither checks nor clips them (it relies on SB3 doing so, hover.py:59-61,
    #     334-341) -- the raw values reach the setpoint and the last four observation columns (oracle-only fixture)
    rngw = np.random.default_rng(5)
    run_scenario(hover, "wild_actions", seed=3, noise=True, n_steps=24, action_fn=lambda k, state: rngw.uniform(-2.5, 2.5, 4))
    rng60 = np.random.default_rng(77)
    run_scenario(hover, "agent_hz60", seed=21, noise=True, n_steps=60, action_fn=lambda k, state: hover_actions(state[None], tgt, rng60, 0.05)[0], agent_hz=60)


if __name__ == "__main__":
    main()
