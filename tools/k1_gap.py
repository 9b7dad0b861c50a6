#!/usr/bin/env python
"""Where does the time between the env-step launches go?  (GPU box.)  Times the same 1 Mi-env hover step
  A  launched from Python, one CUDA event per step (what bench.py does),
  B  launched from Python, events only around the whole loop,
  C  32 steps captured in one CUDA graph and replayed,
  D  step launch only (qx_step_begin; nothing finishes in these steps, the reset-queue launch is pure overhead),
and prints the per-step list of A so that the distribution (not only its mean) is on record.
Usage: python tools/k1_gap.py [envs] [steps]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402
from fpv_drone_rl_agent_b200 import _lib  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 24  # < 32: no mass reset inside the window
dev = torch.device("cuda", 0)
L = _lib.lib()
g = torch.Generator(device="cpu").manual_seed(0)
acts = torch.rand(8, E, 4, generator=g) * 2 - 1
acts[..., :3] *= 0.3
acts[..., 3] = (2 * 0.4952 - 1) + 0.3 * acts[..., 3]
acts = acts.to(dev)
obs = torch.zeros(E, 20, device=dev)
rew = torch.zeros(E, device=dev)
te = torch.zeros(E, dtype=torch.uint8, device=dev)
tr = torch.zeros(E, dtype=torch.uint8, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731


def fresh():
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=0.4952, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
    sim.reset(obs)
    for k in range(3):
        sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    return sim


def timed(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(n):
        fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


out = {}
sim = fresh()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS + 1)]
ev[0].record()
for k in range(STEPS):
    sim.step(acts[k % 8], obs, rew, te, tr)
    ev[k + 1].record()
torch.cuda.synchronize()
per = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(STEPS)]
out["A_python_event_per_step_us"] = sum(per) / len(per)
out["A_per_step"] = [round(x, 1) for x in per]
sim.close()

sim = fresh()
out["B_python_no_events_us"] = timed(lambda k: sim.step(acts[k % 8], obs, rew, te, tr), STEPS)
sim.close()

sim = fresh()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        for k in range(8):
            sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(3):  # 24 steps after the 3 + 8 (capture does not run) warm-up steps
        gr.replay()
    b.record(s)
    torch.cuda.synchronize()
    out["C_graph_us"] = a.elapsed_time(b) * 1e3 / 24
sim.close()

sim = fresh()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def begin_only(k):
    _lib.check(L.qx_step_begin(sim._h, vp(acts[k % 8]), vp(obs), 0, 20, vp(rew), vp(te), vp(tr), None, st))


out["D_step_launch_only_us"] = timed(begin_only, STEPS)
sim.close()
print(json.dumps(out))
