#!/usr/bin/env python
"""Time the hover env step for the kernel variants (GPU box).  Usage: python tools/k1_variants.py [envs] [steps]
Each variant runs in a fresh handle: QX_HOT / QX_LANES / QX_SHAPE / QX_MERGED are read at qx_create.
K1_RATE_SCALE (default 0.3) scales the random roll / pitch / yaw-rate actions: 0.3 tumbles every drone into the floor within
~30 steps (the round-1 bench workload), 0.02 keeps the fleet airborne (SURVEY 8d C2 "so envs stay airborne")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
acts = torch.rand(8, E, 4, generator=g) * 2 - 1
acts[..., :3] *= float(os.environ.get("K1_RATE_SCALE", "0.3"))
acts[..., 3] = (2 * 0.4952 - 1) + 0.3 * acts[..., 3]
acts = acts.to(dev)
obs = torch.zeros(E, 20, device=dev)
rew = torch.zeros(E, device=dev)
te = torch.zeros(E, dtype=torch.uint8, device=dev)
tr = torch.zeros(E, dtype=torch.uint8, device=dev)
variants = [("generic_two_launches", {"QX_HOT": "0"})]
variants += [(f"hot_two_launches_shape{sh}", {"QX_HOT": "1", "QX_LANES": "1", "QX_SHAPE": str(sh), "QX_MERGED": "0"}) for sh in (4,)]
variants += [(f"hot_merged_shape{sh}", {"QX_HOT": "1", "QX_LANES": "1", "QX_SHAPE": str(sh), "QX_MERGED": "1"}) for sh in (3, 4, 5)]
variants += [("hot_merged_2env_packed_shape0", {"QX_HOT": "1", "QX_LANES": "2", "QX_SHAPE": "0", "QX_MERGED": "1"})]
only = os.environ.get("K1_ONLY")
if only:
    variants = [v for v in variants if v[0] in only.split(",")]
import ctypes as C  # noqa: E402
L = pkg._lib.lib()
probe = torch.zeros(4, dtype=torch.int64, device=dev)
res = {}
for name, env in variants:
    os.environ.update(env)
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=0.4952, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
    sim.reset(obs)
    for k in range(5):
        sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS + 1)]
    L.qx_debug_clock_probe(C.c_void_p(probe.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ev[0].record()
    for k in range(STEPS):
        sim.step(acts[k % 8], obs, rew, te, tr)
        ev[k + 1].record()
    L.qx_debug_clock_probe(C.c_void_p(probe.data_ptr() + 16), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    pr = probe.tolist()
    sm_mhz = (pr[2] - pr[0]) / max(pr[3] - pr[1], 1) * 1e3
    raw = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(STEPS)]
    per = sorted(raw)
    res[name] = {"first8_us": sum(raw[:8]) / 8, "last8_us": sum(raw[-8:]) / 8, "mean_us": sum(per) / len(per), "min_us": per[0], "median_us": per[len(per) // 2], "p90_us": per[int(0.9 * len(per))],
                 "frac_of_hbm_roofline_mean": 358.0 * E / (sum(per) / len(per) * 1e-6) / 6538.6e9, "checksum": float(rew.double().sum()), "sm_mhz_avg_over_loop": sm_mhz}
    sim.close()
    print(name, json.dumps(res[name]), flush=True)
