#!/usr/bin/env python
"""PPO rollout side measurements (BASELINE metric M2: PPO samples/s), importable from bench.py or run alone:

    python tools/bench_ppo.py [--envs 131072] [--steps 32] [--task hover]

Reports, for one GPU: full on-device rollout samples/s (env step + reset + VecNormalize + policy forward +
sampling + bootstrap + GAE, CUDA graph), and the stand-alone policy-forward (tensor-core) and GAE kernels against
their rooflines."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLOP_PER_SAMPLE = 77056  # [128,128] pi and vf nets on 20-D obs
GAE_BYTES = 17  # r, V, done in; adv, ret out


def measure(envs: int = 131072, steps: int = 32, reps: int = 5, rank: int = 0, world: int = 1, rollout_only: bool = False,
            task: int = 0) -> dict:
    """task 0 = hover (20-D obs, 4-D action); task 1 = yaw (12-D obs, 1-D action; BASELINE configs[2], rollout only)."""
    import torch

    from fpv_drone_rl_agent_b200 import ppo

    dev = torch.device("cuda", torch.cuda.current_device())
    peaks = {}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks = json.load(open(pp))
    tflops_peak = peaks.get("bf16_tflops", 1590.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    cfg = ppo.PPOConfig(n_envs=envs, n_steps=steps, seed=0, use_cuda_graph=True,
                        chain_obs_stats=os.environ.get("PPO_CHAIN_OBS_STATS", "1") != "0",  # (A/B switches for tools runs)
                        bootstrap_first={"0": False, "1": True}.get(os.environ.get("PPO_BOOTSTRAP_FIRST", ""), None))
    tr = ppo.PPOTrainer(cfg, device=dev, rank=rank, world=world, task=task)
    ro = tr.rollout

    def timed(fn, reps=reps):
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return statistics.median(ts)

    ro.collect(); ro.collect()
    torch.cuda.synchronize()
    ms = timed(ro.collect)
    out = {"task": "yaw" if task else "hover", "envs": envs, "n_steps": steps,
           "rollout": {"samples_per_s": envs * steps / (ms * 1e-3), "ms_per_rollout": ms, "launches_per_rollout": ro.launches_per_rollout,
                       "what": "env step + reset + obs/reward normalisation + policy forward + sampling + time-limit bootstrap, x n_steps, + GAE; one CUDA graph replay"}}
    if task == 0:
        out["update"] = measure_update(tr, world)
    if rollout_only or task != 0:
        tr.sim.close()
        return out
    # stand-alone policy forward over a batch larger than L2 is pointless (weights are tiny; activations stay on chip):
    # use the rollout's own batch, all outputs on
    n = envs
    obs = torch.randn(n, 20, device=dev)
    acts = torch.zeros(n, 4, device=dev); eacts = torch.zeros(n, 4, device=dev); vals = torch.zeros(n, device=dev)
    logp = torch.zeros(n, device=dev); on = torch.zeros(n, 20, device=dev)
    f = lambda: ppo.policy_forward(tr.packed, obs, obs_stats=ro.obs_stats, seed=1, step=3, actions=acts, env_actions=eacts, values=vals,  # noqa: E731
                                   log_probs=logp, obs_norm=on)
    f(); torch.cuda.synchronize()
    ms = timed(f, reps=9)
    tf = n * FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12
    out["policy_forward"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms,
                             "roofline": {"bound": "tensor", "achieved": tf, "peak": tflops_peak, "unit": "TFLOP/s", "frac": tf / tflops_peak,
                                          "note": "77 056 algorithmic FLOP per sample; the kernel is bounded by the 512 tanh (MUFU) per sample, see DESIGN.md K2"}}
    T = steps
    rew = torch.randn(T, n, device=dev); val = torch.randn(T, n, device=dev); dn = torch.zeros(T, n, dtype=torch.uint8, device=dev)
    last = torch.randn(n, device=dev); adv = torch.zeros(T, n, device=dev); ret = torch.zeros(T, n, device=dev)
    g = lambda: ppo.gae(rew, val, dn, last, 0.99, 0.95, adv, ret)  # noqa: E731
    g(); torch.cuda.synchronize()
    ms = timed(g, reps=9)
    gb = T * n * GAE_BYTES / (ms * 1e-3) / 1e9
    out["gae"] = {"ms": ms, "roofline": {"bound": "hbm", "achieved": gb, "peak": hbm_peak, "unit": "GB/s", "frac": gb / hbm_peak}}
    tr.sim.close()
    return out


def measure_update(tr, world: int = 1, batch_size: int = 262144, epochs: int = 2) -> dict:
    """K5: the PPO update on the rollout just collected -- optimiser steps of `batch_size` rows (forward + loss + backward on
    the tensor cores, gradient reduction, all-reduce when world > 1, clip + Adam + re-pack), one CUDA-graph replay per epoch.
    Reports ms per optimiser step, trained sample-passes/s, and -- with several ranks -- the all-reduce of the flat gradient
    timed alone (the only collective of the path)."""
    import time

    import torch
    import torch.distributed as dist

    cfg, ro = tr.cfg, tr.rollout
    old = (cfg.batch_size, cfg.n_epochs)
    cfg.batch_size, cfg.n_epochs = min(batch_size, ro.T * ro.n), epochs
    tr._epoch_graph = None
    tr.update()  # builds the graph
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    st = tr.update()
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = s.elapsed_time(e)
    steps = st["optimizer_steps"]
    rows = steps * (cfg.batch_size // 128) * 128
    out = {"batch_size_per_rank": cfg.batch_size, "epochs": epochs, "optimizer_steps": steps, "update_ms_per_step": ms / max(steps, 1),
           "update_ms_total": ms, "wall_ms_total": 1e3 * wall, "sample_passes_per_s_per_gpu": rows / (ms * 1e-3),
           "flop_per_sample_pass": 3 * FLOP_PER_SAMPLE, "tflops_algorithmic": rows * 3 * FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12,
           "kernels": "ppo_upd::update_prepare / update_fwdbwd (tcgen05) / update_reduce / update_adam", "n_params": tr.fused.n_params}
    if world > 1:
        g = tr.fused.grad
        for _ in range(5):
            dist.all_reduce(g)
        torch.cuda.synchronize()
        dist.barrier()
        s.record()
        for _ in range(50):
            dist.all_reduce(g)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 50 * 1e3], device=g.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["allreduce_us"] = float(t.item())
        out["allreduce_bytes"] = g.numel() * 4
        g.zero_()
    cfg.batch_size, cfg.n_epochs = old
    tr._epoch_graph = None
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--task", choices=["hover", "yaw"], default="hover")
    a = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    print(json.dumps(measure(a.envs, a.steps, task=1 if a.task == "yaw" else 0)))
