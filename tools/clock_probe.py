#!/usr/bin/env python
"""Where does the gap between the ncu kernel durations and the event-timed step go?  (GPU box only.)

Runs the hover step loop of bench.py for a few hundred steps and prints (a) per-step event times at the start / end of the
run, (b) the SM clock sampled through NVML every ~2 ms while the loop runs, (c) the same loop captured into one CUDA graph,
(d) the loop with no events between the steps.  Usage: python tools/clock_probe.py [envs] [steps]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda", 0)
cfg = pkg.default_config()
cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=0.4952, auto_reset=1, noise=1)
sim = pkg.QuadXSim(E, cfg, seed=1234, device=dev)
g = torch.Generator(device="cpu").manual_seed(0)
acts = torch.rand(8, E, 4, generator=g) * 2 - 1
acts[..., :3] *= 0.3
acts[..., 3] = (2 * 0.4952 - 1) + 0.3 * acts[..., 3]
acts = acts.to(dev)
obs = torch.zeros(E, 20, device=dev)
rew = torch.zeros(E, device=dev)
te = torch.zeros(E, dtype=torch.uint8, device=dev)
tr = torch.zeros(E, dtype=torch.uint8, device=dev)
sim.reset(obs)
for k in range(5):
    sim.step(acts[k % 8], obs, rew, te, tr)
torch.cuda.synchronize()

import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = False


def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.002)


res = {"envs": E, "steps": STEPS}
th = threading.Thread(target=sampler)
th.start()
time.sleep(0.05)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS + 1)]
t0 = time.perf_counter()
ev[0].record()
for k in range(STEPS):
    sim.step(acts[k % 8], obs, rew, te, tr)
    ev[k + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
per = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(STEPS)]
inside = [(c, p) for (t, c, p) in samples if t0 <= t <= t1]
res["events_between_steps"] = {
    "mean_us": sum(per) / len(per), "first10_us": [round(x, 1) for x in per[:10]], "last10_us": [round(x, 1) for x in per[-10:]],
    "min_us": min(per), "sorted_deciles_us": [round(sorted(per)[int(q * (len(per) - 1) / 10)], 1) for q in range(11)],
    "sm_mhz_samples": [c for c, _ in inside][:: max(1, len(inside) // 40)], "power_w": [round(p) for _, p in inside][:: max(1, len(inside) // 40)],
}
# (d) no events between steps
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(STEPS):
    sim.step(acts[k % 8], obs, rew, te, tr)
b.record()
torch.cuda.synchronize()
res["no_events_mean_us"] = a.elapsed_time(b) * 1e3 / STEPS
# (c) one CUDA graph of 32 steps
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        for k in range(32):
            sim.step(acts[k % 8], obs, rew, te, tr)
    gr.replay()
    torch.cuda.synchronize()
    a.record(s)
    for _ in range(max(1, STEPS // 32)):
        gr.replay()
    b.record(s)
torch.cuda.synchronize()
res["graph32_mean_us"] = a.elapsed_time(b) * 1e3 / (32 * max(1, STEPS // 32))
stop = True
th.join()
print(json.dumps(res))
