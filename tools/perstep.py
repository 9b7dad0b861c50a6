#!/usr/bin/env python
"""Per-step anatomy of the 1 Mi-env hover step (GPU box): time of the step launch and of the reset-queue launch
(events between qx_step_begin and qx_step_end), and how many envs finished, for every step of a window.
Workloads: tumble (rate actions x0.3), gentle (x0.02), still (zero rate actions, hover thrust, noise off).
Usage: python tools/perstep.py [envs] [steps]   (env: QX_* kernel variables as usual)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402

HOVER_THR = (0.1 * 9.81 / 4.0) ** 0.5
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 72
dev = torch.device("cuda", 0)
obs = torch.zeros(E, 20, device=dev)
rew = torch.zeros(E, device=dev)
te = torch.zeros(E, dtype=torch.uint8, device=dev)
tr = torch.zeros(E, dtype=torch.uint8, device=dev)
for name, rate, thr_amp, noise in (("tumble", 0.3, 0.3, 1), ("gentle", 0.02, 0.3, 1), ("still", 0.0, 0.0, 0)):
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=noise)
    sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    acts = torch.rand(8, E, 4, generator=g) * 2 - 1
    acts[..., :3] *= rate
    acts[..., 3] = (2 * HOVER_THR - 1) + thr_amp * acts[..., 3]
    acts = acts.to(dev)
    sim.reset(obs)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    dones = []
    for k in range(K):
        ev[k][0].record()
        pkg._lib.check(sim.lib.qx_step_begin(sim._h, acts[k % 8].data_ptr(), obs.data_ptr(), 0, 20, rew.data_ptr(), te.data_ptr(), tr.data_ptr(), None,
                                             torch.cuda.current_stream().cuda_stream))
        ev[k][1].record()
        pkg._lib.check(sim.lib.qx_step_end(sim._h, obs.data_ptr(), 0, 20, torch.cuda.current_stream().cuda_stream))
        ev[k][2].record()
        dones.append((te | tr).sum())
    torch.cuda.synchronize()
    rows = [(round(ev[k][0].elapsed_time(ev[k][1]) * 1e3, 1), round(ev[k][1].elapsed_time(ev[k][2]) * 1e3, 1), int(dones[k])) for k in range(K)]
    print(name, json.dumps({"step_us|reset_us|n_done per step": rows}))
    sim.close()
