#!/usr/bin/env python
"""Whole-step time of qx_step (step launch + reset-queue launch, one CUDA event per step) at 1 Mi envs for the current QX_*
settings: median of steps 3..14 (nothing finishes) and mean of steps 5..24 (bench.py's window at the driver's flags)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402

HOVER_THR = (0.1 * 9.81 / 4.0) ** 0.5
E, K = 1 << 20, 30
dev = torch.device("cuda", 0)
obs = torch.zeros(E, 20, device=dev); rew = torch.zeros(E, device=dev)
te = torch.zeros(E, dtype=torch.uint8, device=dev); tr = torch.zeros(E, dtype=torch.uint8, device=dev)
out = {}
for name, rate in (("tumble", 0.3), ("gentle", 0.02)):
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    acts = torch.rand(8, E, 4, generator=g) * 2 - 1
    acts[..., :3] *= rate
    acts[..., 3] = (2 * HOVER_THR - 1) + 0.3 * acts[..., 3]
    acts = acts.to(dev)
    res = []
    for rep in range(3):
        sim.reset(obs)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record()
        for k in range(K):
            sim.step(acts[k % 8], obs, rew, te, tr)
            ev[k + 1].record()
        torch.cuda.synchronize()
        per = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(K)]
        res.append((round(statistics.median(per[3:15]), 1), round(statistics.mean(per[5:25]), 1)))
    # the same window without an event between the steps (an event record between two kernels ends a programmatic launch chain)
    tot = []
    for rep in range(3):
        sim.reset(obs)
        for k in range(5):
            sim.step(acts[k % 8], obs, rew, te, tr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(5, 25):
            sim.step(acts[k % 8], obs, rew, te, tr)
        e1.record()
        torch.cuda.synchronize()
        tot.append(round(e0.elapsed_time(e1) * 1e3 / 20, 1))
    out[name] = res
    out[name + "_steps5to24_no_inner_events"] = tot
    sim.close()
print(json.dumps({"QX_PDL": os.environ.get("QX_PDL", "1"), **out}))
