#!/usr/bin/env python
"""Time qx_step_host_ex (bf16 observations, pinned host buffers) at 1 Mi envs for the current QX_HOST_* settings (GPU box)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402
from fpv_drone_rl_agent_b200 import _lib  # noqa: E402

E = 1 << 20
HOVER_THR = (0.1 * 9.81 / 4.0) ** 0.5
dev = torch.device("cuda", 0)
cfg = pkg.default_config()
cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
acts = torch.rand(8, E, 4) * 2 - 1
acts[..., :3] *= 0.3
acts[..., 3] = (2 * HOVER_THR - 1) + 0.3 * acts[..., 3]
acts = acts.pin_memory()
h_obs = torch.zeros(E, 20, dtype=torch.bfloat16).pin_memory()
h_rew = torch.zeros(E).pin_memory()
h_te = torch.zeros(E, dtype=torch.uint8).pin_memory()
h_tr = torch.zeros(E, dtype=torch.uint8).pin_memory()
obs = torch.zeros(E, 20, device=dev)
sim.reset(obs)
L = _lib.lib()
vp = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
for k in range(3):
    _lib.check(L.qx_step_host_ex(sim._h, vp(acts[k % 8]), vp(h_obs), 1, vp(h_rew), vp(h_te), vp(h_tr), None))
ts = []
for k in range(20):
    t0 = time.perf_counter()
    _lib.check(L.qx_step_host_ex(sim._h, vp(acts[k % 8]), vp(h_obs), 1, vp(h_rew), vp(h_te), vp(h_tr), None))
    ts.append(time.perf_counter() - t0)
ts.sort()
print(json.dumps({"chunks": os.environ.get("QX_HOST_CHUNKS", "8"), "one_d2h": os.environ.get("QX_HOST_ONE_D2H", "0"),
                  "ms_median": round(ts[10] * 1e3, 4), "ms_min": round(ts[0] * 1e3, 4), "env_steps_per_s_median": round(E / ts[10] / 1e6, 1)}))
