#!/usr/bin/env python
"""bench.py -- env-steps/s of the QuadX hover step (physics + reward + obs) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]

One "step" = one pass of the hot path (QuadXHoverEnv.step for every env: 12
physics sub-steps, 6 control updates, reward, termination, auto-reset, obs)
over one batch of E envs per GPU.  Prints ONE JSON line (rank 0).  See
DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ENV_STEP = 358  # 128+128 state, 16 action, 80 obs, 4 reward, 2 flags
ACT_BYTES, OUT_BYTES, OUT_BYTES_BF16 = 16, 80 + 4 + 1 + 1, 40 + 4 + 1 + 1
HOVER_THR = (0.1 * 9.81 / 4.0) ** 0.5
METRIC = "env-steps/sec (physics+reward+obs), QuadX hover"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-envs", type=int, default=262144, help="envs of the bounded CPU sample (one CPU step = this many envs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small", action="store_true", help="skip the configs[1] (4096 envs) side measurements")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO rollout side measurements (metric M2)")
    ap.add_argument("--no-train", action="store_true", help="skip the bounded PPO training-to-target leg (SURVEY C5)")
    return ap.parse_args()


def workload_config(envs_per_gpu: int, n_gpus: int) -> dict:
    return {
        "workload": "hover (hover.py QuadXHoverEnv), %d envs/GPU x %d GPU, physics+reward+obs+auto-reset step" % (envs_per_gpu, n_gpus),
        "baseline_config": "configs[3] shard size generalised to N=1 (1 Mi envs on one GPU); configs[1] (4096 envs) reported under 'hover_4096'",
        "envs_per_gpu": envs_per_gpu,
        "n_envs_total": envs_per_gpu * n_gpus,
        "sub_steps_per_env_step": 12,
        "spawn": "pos (0,0,1) level at rest, motors at hover throttle; reference reset protocol (10 idle Aviary.step) on auto-reset",
        "actions": "U(-1,1)*0.3 rates, thrust a3 = hover + 0.3*U(-1,1); 8 pre-generated device batches cycled",
        "motor_noise": "2 % (cf2x.yaml:5), Philox seed 1234",
        "l2_policy": "working set per step (state 176 B r+w + outputs, x envs) exceeds the 126 MB L2 when envs/GPU >= 2^19; no flush needed at the default size",
        "sharding": "contiguous env ranges, Philox key = global env id, no data-path collective",
    }


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU arm
def cpu_run(n_envs: int, steps: int, warmup: int, budget_s: float = 25.0) -> dict:
    """Time the CPU oracle (port of the reference path) on a bounded sample of the
    same workload.  Uses the threaded C port when it is built, else numpy."""
    import numpy as np

    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (8, n_envs, 4))
    acts[..., :3] *= 0.3
    acts[..., 3] = (2 * HOVER_THR - 1) + 0.3 * acts[..., 3]
    try:
        from oracle import c_oracle

        sim = c_oracle.COracle(n_envs, seed=1234, start_pos=(0, 0, 1.0), spawn_throttle=HOVER_THR)
        kind_detail, cores = "oracle/quadx_oracle.c (float64, all host threads)", sim.threads
        step = lambda a: sim.step(a)  # noqa: E731
        sim.reset()
    except Exception:
        from oracle.hover_oracle import HoverConfig, HoverVecOracle

        sim = HoverVecOracle(n_envs, cfg=HoverConfig(start_pos=(0, 0, 1.0), spawn_throttle=HOVER_THR), seed=1234, noise=True)
        kind_detail, cores = "oracle/hover_oracle.py (numpy float64, vectorised over envs)", 1
        step = lambda a: sim.step(a)  # noqa: E731
        sim.reset()
    for k in range(warmup):
        step(acts[k % 8])
    t0 = time.perf_counter()
    done = 0
    for k in range(steps):
        step(acts[k % 8])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n_envs * done / dt, "unit": UNIT, "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{n_envs} envs x {done} steps of the same hover workload, {kind_detail}", "seconds": dt, "steps_done": done,
            "ms_per_step": 1e3 * dt / done}


def cpu_extras() -> dict:
    """The other CPU yardsticks of BASELINE.md section 3 (each ~2 s): numpy oracle with one env in a Python loop (the closest
    analogue of one PyFlyt env process, minus rendering), numpy oracle vectorised over 4096 envs, the C port on one thread."""
    import numpy as np

    from oracle.hover_oracle import HoverConfig, HoverVecOracle

    out = {}
    for name, n, budget in (("numpy_1env_python_loop", 1, 2.0), ("numpy_vectorised_4096_envs_1core", 4096, 3.0)):
        sim = HoverVecOracle(n, cfg=HoverConfig(start_pos=(0, 0, 1.0), spawn_throttle=HOVER_THR), seed=1234, noise=True)
        sim.reset()
        a = np.zeros((n, 4)); a[:, 3] = 2 * HOVER_THR - 1
        sim.step(a)
        t0, k = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget:
            sim.step(a); k += 1
        out[name] = n * k / (time.perf_counter() - t0)
    try:
        from oracle import c_oracle

        os.environ["ORC_THREADS"] = "1"
        sim = c_oracle.COracle(16384, seed=1234, start_pos=(0, 0, 1.0), spawn_throttle=HOVER_THR)
        sim.reset()
        a = np.zeros((16384, 4)); a[:, 3] = 2 * HOVER_THR - 1
        sim.step(a)
        t0, k = time.perf_counter(), 0
        while time.perf_counter() - t0 < 2.0:
            sim.step(a); k += 1
        out["c_port_1_thread"] = 16384 * k / (time.perf_counter() - t0)
    except Exception as exc:  # noqa: BLE001
        out["c_port_1_thread"] = f"unavailable: {exc}"
    finally:
        os.environ.pop("ORC_THREADS", None)
    out["reference_logged_tensorboard_fps"] = "832-1003 env-steps/s on 16 SubprocVecEnv processes (author's PC; SAC + PyBullet + 6 CPU renders per step) -- quoted, different machine"
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_run(args.cpu_envs, args.steps, args.warmup, budget_s=150.0)
    cfg = workload_config(args.cpu_envs, 1)
    cfg["workload"] = ("hover (hover.py QuadXHoverEnv), CPU sample of the same workload: %d envs x %d steps on %d host threads "
                       "(the GPU arm runs %d envs/GPU x %d GPU)" % (args.cpu_envs, r["steps_done"], r["cores"], args.envs, args.gpus))
    cfg["gpu_arm_envs_per_gpu"] = args.envs
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps_done"],
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "PyFlyt/pybullet are not installable here (no network); the reference arm is the CPU oracle port of the same path "
                "(oracle/quadx_oracle.c, float64, all host threads); env-steps/s does not depend on the batch size on the CPU",
    }
    print(json.dumps(line))


def hover_leg(pkg, dev, envs: int, rank: int, world: int, steps: int, warmup: int, workload: str) -> dict:
    """One more timing of the hover step on a fresh handle: `workload` = "tumble" (the headline's action distribution) or
    "airborne" (zero rate commands, hover thrust, motor noise on: the drones stay in the air for the whole window, the regime a
    trained policy keeps the fleet in -- no env touches the floor, nothing finishes, the reset queue stays empty)."""
    import torch
    import torch.distributed as dist

    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(envs, cfg, seed=1234, env_id0=rank * envs, device=dev)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    acts = torch.rand(8, envs, 4, generator=g) * 2 - 1
    if workload == "airborne":
        acts[..., :3] = 0.0
        acts[..., 3] = 2 * HOVER_THR - 1
    else:
        acts[..., :3] *= 0.3
        acts[..., 3] = (2 * HOVER_THR - 1) + 0.3 * acts[..., 3]
    acts = acts.to(dev)
    obs = torch.zeros(envs, 20, device=dev); rew = torch.zeros(envs, device=dev)
    te = torch.zeros(envs, dtype=torch.uint8, device=dev); tr = torch.zeros(envs, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if envs * 454 < (160 << 20) else None  # state + outputs fit in L2: flush between steps
    sim.reset(obs)
    for k in range(max(warmup, 3)):
        sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if flush is None:  # one event pair around the window: an event record between two steps would end the programmatic launch chain
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            sim.step(acts[k % 8], obs, rew, te, tr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()
            ev[k][0].record()
            sim.step(acts[k % 8], obs, rew, te, tr)
            ev[k][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sim.close()
    return {"workload": workload, "envs_per_gpu": envs, "n_gpus": world, "steps": steps, "ms_per_step": ms / steps,
            "env_steps_per_s": envs * world * steps / (ms * 1e-3), "l2": "flushed between steps (256 MB memset, outside the timed events)" if flush is not None else "working set exceeds L2"}


def train_leg(dev, rank: int, world: int, envs_total: int = 16384, budget_s: float = 40.0, target_len: float = 400.0, target_rew: float = 181.0) -> dict:
    """SURVEY C5 as a bounded bench leg: PPO hover training from scratch on the reference's reset protocol until
    rollout/ep_len_mean >= 402-ish (the cap the reference's runs saturate at) and rollout/ep_rew_mean >= 0.5 * 402 * 0.9 = 181,
    or until the time budget is spent.  Rollout, update (hand-written kernels) and, with several ranks, the gradient all-reduce.
    STRONG scaling: the same problem at every N -- 16 384 envs and a global minibatch of 32 768 rows, split over the ranks -- because
    what limits time-to-target is the number of trust-region iterations, not the data per iteration (measured: the per-rank config
    kept fixed on 2 GPUs needs more wall clock to the reward target than one GPU, profiles/train_hover_r2.md)."""
    import torch
    import torch.distributed as dist

    from fpv_drone_rl_agent_b200 import ppo

    envs = max(128, envs_total // world)

    # hyper-parameters: tools/train_sweep.py "fs_ent" (profiles/train_hover_r2.md): lr 3e-4 (train_hover.py:56) with a linear decay, clip 0.2,
    # gamma 0.995 / lambda 0.97, small initial action noise; an entropy bonus of 0.003 and a floor of -4.5 under log_std for the first
    # 300 iterations (released to -7 by iteration 900) keep the policy exploring until it has left the ~165 plateau: 5 of 5 seeds reach
    # both targets within 5-11 s (without them 1 of 5 within 45 s)
    cfg = ppo.PPOConfig(n_envs=envs, n_steps=64, n_epochs=4, batch_size=max(128, 32768 // world), learning_rate=3e-4, seed=1, target_kl=0.02, log_std_init=-1.6,
                        gamma=0.995, gae_lambda=0.97, ent_coef=0.003, lr_final_frac=0.05, lr_anneal_iters=3000,
                        log_std_min=-4.5, log_std_min_final=-7.0, log_std_min_iters=(300, 900))
    tr = ppo.PPOTrainer(cfg, device=dev, rank=rank, world=world)
    tr.learn_iteration()  # graph capture / one-time setup outside the clock (its samples still count as training)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_len = t_rew = None
    best_rew, it, out = -1e9, 0, {}
    while True:
        out = tr.learn_iteration()
        it += 1
        now = time.perf_counter() - t0
        if out["ep_len_mean"] == out["ep_len_mean"]:
            best_rew = max(best_rew, out["ep_rew_mean"])
            if t_len is None and out["ep_len_mean"] >= target_len:
                t_len = now
            if t_rew is None and out["ep_len_mean"] >= target_len and out["ep_rew_mean"] >= target_rew:
                t_rew = now
        stop = torch.tensor([1.0 if (t_rew is not None or now > budget_s) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(stop, op=dist.ReduceOp.MAX)
        if float(stop) > 0:
            break
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    res = {"scaling": "strong (16 384 envs and a 32 768-row global minibatch in total, split over the ranks)", "envs_per_gpu": envs, "n_gpus": world, "n_steps": cfg.n_steps, "n_epochs": cfg.n_epochs, "batch_size_per_rank": cfg.batch_size, "iterations": it,
           "env_steps": int(out["timesteps"]), "wall_s": wall, "env_steps_per_s_incl_update": (out["timesteps"] - envs * cfg.n_steps * world) / wall,
           "target_ep_len": target_len, "target_ep_rew": target_rew, "time_to_ep_len_s": t_len, "time_to_ep_len_and_rew_s": t_rew,
           "final_ep_len_mean": out["ep_len_mean"], "final_ep_rew_mean": out["ep_rew_mean"], "best_ep_rew_mean": best_rew, "budget_s": budget_s,
           "reference": "train_hover.py logs: ep_len_mean reaches the 402 cap after ~230-250 k steps = 275-300 s on 16 CPU processes (SURVEY 6)"}
    tr.sim.close()
    return res


# ---------------------------------------------------------------------------- ours
def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import fpv_drone_rl_agent_b200 as pkg
    from fpv_drone_rl_agent_b200 import _lib

    E = args.envs
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(E, cfg, seed=1234, env_id0=rank * E, device=dev)
    g = torch.Generator(device="cpu").manual_seed(rank)
    acts = torch.rand(8, E, 4, generator=g) * 2 - 1
    acts[..., :3] *= 0.3
    acts[..., 3] = (2 * HOVER_THR - 1) + 0.3 * acts[..., 3]
    acts_pinned = acts.pin_memory()
    acts = acts.to(dev)
    obs = torch.zeros(E, 20, device=dev)
    rew = torch.zeros(E, device=dev)
    te = torch.zeros(E, dtype=torch.uint8, device=dev)
    tr = torch.zeros(E, dtype=torch.uint8, device=dev)
    sim.reset(obs)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)  # (before the warm-up, not after it: a quarter of a second of idling lets the GPU leave its working clocks)
    if world > 1:
        dist.barrier()
    for k in range(args.warmup):
        sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    if os.environ.get("BENCH_IDLE_BEFORE_TIMING"):  # diagnostic: the round-1 / early round-2 order (idle gap between warm-up and timing)
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.lib().qx_launch_count()
    # the timed region: K steps between two events on the launching stream.  No event between the steps: a tight loop of qx_step is
    # a chain of programmatic dependent launches (step -> reset queue -> next step), which an event record in between would cut
    # (tools/step_ab.py: 139.0 us per step against 142.9 with one event per step); the per-step times of the line come from a
    # second, diagnostic pass over the same window below
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        sim.step(acts[k % 8], obs, rew, te, tr)
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    launches = _lib.lib().qx_launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)
    # diagnostic pass (not part of `value`): the same window again from a fresh reset, one event per step
    sim.reset(obs)
    for k in range(args.warmup):
        sim.step(acts[k % 8], obs, rew, te, tr)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        sim.step(acts[k % 8], obs, rew, te, tr)
        ev[k + 1].record()
    torch.cuda.synchronize()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = E * world * args.steps / (max_ms * 1e-3)

    # ---- e2e: the C-ABI host-buffer call, H2D + kernel + D2H inside the timed region.  The call is bound by the device-to-host
    # copy, so it is measured with both wire formats of the observation: bf16 (qx_step_host_ex, what a bf16 policy input needs;
    # the headline `e2e`) and f32 (qx_step_host, `e2e_f32_obs`).  Physics, reward and flags are the same f32 arithmetic in both.
    import ctypes as C

    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)  # pinned buffers are allocated on the host NUMA node the GPU hangs off
    acts_pinned = acts_pinned.clone().pin_memory() if numa is not None else acts_pinned
    h_obs = torch.zeros(E, 20).pin_memory()
    h_obs16 = torch.zeros(E, 20, dtype=torch.bfloat16).pin_memory()
    h_rew = torch.zeros(E).pin_memory()
    h_te = torch.zeros(E, dtype=torch.uint8).pin_memory()
    h_tr = torch.zeros(E, dtype=torch.uint8).pin_memory()

    L = _lib.lib()
    vp = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_run(bf16: bool) -> float:
        def host_step(k):
            if bf16:
                _lib.check(L.qx_step_host_ex(sim._h, vp(acts_pinned[k % 8]), vp(h_obs16), 1, vp(h_rew), vp(h_te), vp(h_tr), None))
            else:
                _lib.check(L.qx_step_host(sim._h, vp(acts_pinned[k % 8]), vp(h_obs), vp(h_rew), vp(h_te), vp(h_tr), None))

        for k in range(3):
            host_step(k)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0 = time.perf_counter()
        for k in range(e2e_steps):
            host_step(k)
        e1 = time.perf_counter()
        te2e = torch.tensor([e1 - e0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        return E * world * e2e_steps / float(te2e.item())

    e2e_f32 = e2e_run(False)
    e2e_value = e2e_run(True)
    os.sched_setaffinity(0, affinity0)  # the CPU legs below use every host thread again

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern_ms = total_ms / args.steps  # this rank's timed region (the line's ms_per_step is the slowest rank's)
        achieved = E * ALG_BYTES_PER_ENV_STEP / (kern_ms * 1e-3) / 1e9
        traffic, compute_side, traffic_stale = None, None, None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            import hashlib

            hsh = hashlib.sha256()
            for f in ("qx_kernels.cu", "qx_model.cuh", "qx_lanes.cuh", "qx_ref_constants.cuh"):
                hsh.update(open(os.path.join(ROOT, "fpv-drone-rl-agent_b200", "csrc", f), "rb").read())
            traffic_stale = tj.get("kernel_sources_sha16") != hsh.hexdigest()[:16]  # the capture was taken from other kernel sources
            if tj.get("envs") == E:
                traffic = tj.get("dram_bytes_per_launch")
                if "thread_instructions_per_env_step" in tj:
                    # the kernel sits on the compute side of the ridge: at 4 warp-instructions per clock per SM the instruction
                    # count alone caps it below the HBM roofline (numbers from the committed ncu capture, not measured live)
                    ipe = tj["thread_instructions_per_env_step"]
                    cap = 148 * 4 * 32 * 1.965e9 / ipe
                    compute_side = {"thread_instructions_per_env_step": ipe, "issue_active_pct_ncu": tj.get("issue_active_pct"),
                                    "issue_bound_env_steps_per_s": cap, "issue_bound_as_frac_of_hbm_roofline": cap * ALG_BYTES_PER_ENV_STEP / (peak * 1e9),
                                    "source": tj.get("source"), "capture_is_of_these_sources": not traffic_stale}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(E, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * ACT_BYTES, "d2h_bytes_per_step": E * OUT_BYTES_BF16,
                    "steps": e2e_steps, "api": "qx_step_host_ex (C-ABI, pinned host buffers, synchronous; f32 actions in, bf16 observations + f32 reward + u8 terminated / truncated out)",
                    "host_numa_node": numa},
            "e2e_f32_obs": {"value": e2e_f32, "unit": UNIT, "h2d_bytes_per_step": E * ACT_BYTES, "d2h_bytes_per_step": E * OUT_BYTES,
                            "steps": e2e_steps, "api": "qx_step_host (same call with f32 observations)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "kernels": ("reference-constant instantiation (model constants of the reference's own parameter set as literals)"
                        if sim.lib.qx_uses_reference_constants(sim._h) else "generic (every constant read from the config)"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_capture_is_of_these_sources": (None if traffic_stale is None else not traffic_stale),
                         "compute_side": compute_side,
                         "kernel": "qx::quadx_step_hot_kernel<REF, SHAPE 4, S1> (one env per thread, paired FFMA2 / FMUL2 / FADD2) + qx::quadx_reset_hot_kernel, both launches inside the step time", "alg_bytes_per_env_step": ALG_BYTES_PER_ENV_STEP,
                         "kernel_ms": kern_ms, "kernel_ms_min": min(per), "peak_source": peak_src},
        }
        line["roofline"]["step_ms_first8_min"] = min(per[:8]) if len(per) >= 8 else None
        line["roofline"]["step_ms_each"] = [round(x, 4) for x in per[:64]]
        line["roofline"]["step_ms_each_note"] = "diagnostic second pass over the same window with one event per step (costs ~3 us per step: the event record ends the programmatic launch chain); not what `value` was timed on"
        if world == 1 and not args.no_small:
            line["hover_4096"] = small_config(pkg, dev)
        if world == 1 and not args.no_small:
            line["yaw_same_batch"] = yaw_side(pkg, dev, E)
        if world == 1 and not args.no_ppo:
            sim.close()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_ppo

            line["ppo"] = bench_ppo.measure(131072, 32)
            # BASELINE configs[2]: yaw task, 65 536 envs, full on-device rollout (policy forward + sampling + GAE), T = 128
            line["ppo"]["yaw_65536"] = bench_ppo.measure(65536, 128, reps=3, task=1)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_run(args.cpu_envs, 150, 2, budget_s=12.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus")}
            line["cpu_baseline"]["others_env_steps_per_s"] = cpu_extras()
    sim.close()
    if not args.no_small:
        # the same step in the regime a trained policy keeps the fleet in (nobody on the floor, nothing to reset), and BASELINE
        # configs[3] as written: 1 Mi envs in total, split over the GPUs (strong scaling; the per-GPU shard fits in L2 from N = 2)
        air = hover_leg(pkg, dev, E, rank, world, 32, args.warmup, "airborne")
        strong = hover_leg(pkg, dev, max((1 << 20) // world, 4096), rank, world, 32, args.warmup, "tumble")
        if rank == 0:
            air["frac_of_hbm_roofline"] = air["env_steps_per_s"] / world * ALG_BYTES_PER_ENV_STEP / (peak * 1e9)
            line["hover_airborne"] = air
            strong["frac_of_hbm_roofline_per_gpu"] = strong["env_steps_per_s"] / world * ALG_BYTES_PER_ENV_STEP / (peak * 1e9)
            line["strong_1Mi"] = strong
    if not args.no_train:
        tl = train_leg(dev, rank, world)
        if rank == 0:
            line["train"] = tl
    if world > 1 and not args.no_ppo:
        # metric M2 at N GPUs: every rank collects its own rollouts (env shard + policy replica, no collective in the
        # rollout); aggregate samples/s = total samples / slowest rank
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_ppo

        m = bench_ppo.measure(131072, 32, rank=rank, world=world, rollout_only=True)
        t = torch.tensor([m["rollout"]["ms_per_rollout"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            m["rollout"]["ms_per_rollout"] = float(t.item())
            m["rollout"]["samples_per_s"] = world * 131072 * 32 / (float(t.item()) * 1e-3)
            m["rollout"]["what"] += "; aggregate over all ranks, slowest rank's time"
            line["ppo"] = m
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))



def bind_to_gpu_numa_node(local: int):
    """Run this process on the CPUs of the NUMA node GPU `local` is attached to, so that the pinned host buffers it allocates
    afterwards (first touch) are local to that GPU's PCIe root.  Returns the node, or None where sysfs does not say."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def yaw_side(pkg, dev, envs: int) -> dict:
    """The yaw task (yaw.py: one Aviary.step = 2 physics sub-steps per env step) at the same batch size: with 6x less
    arithmetic per step the same kernel is HBM-bound, which shows what the memory path alone sustains.
    Algorithmic bytes per env-step: 128 + 128 state, 4 action, 48 obs, 4 reward, 2 flags = 314 B."""
    import statistics as st

    import torch

    env = pkg.QuadXYawVecEnv(envs, seed=1234, device=dev)
    env.reset()
    g = torch.Generator(device="cpu").manual_seed(0)
    acts = (torch.rand(8, envs, 1, generator=g) * 2 - 1).to(dev)
    obs, rew, te, tr = env.obs, env.rewards, env.terminated, env.truncated
    for k in range(5):
        env.sim.step(acts[k % 8], obs, rew, te, tr)
    K = 32
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for k in range(K):
        env.sim.step(acts[k % 8], obs, rew, te, tr)
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = st.mean(ev[k].elapsed_time(ev[k + 1]) for k in range(K))
    env.close()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    gbs = envs * 314 / (ms * 1e-3) / 1e9
    moved = envs * (176 * 2 + 4 + 48 + 4 + 2) / (ms * 1e-3) / 1e9
    return {"envs": envs, "env_steps_per_s": envs / (ms * 1e-3), "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "alg_bytes_per_env_step": 314,
                         "moved_gbs": moved, "moved_frac": moved / peak}}


def small_config(pkg, dev) -> dict:
    """BASELINE.json configs[1]: 4096 envs on one GPU -- latency-bound; reported
    for per-step launches, a CUDA graph of 64 steps and step_k (K = 64)."""
    import torch

    n, K = 4096, 64
    cfg = pkg.default_config()
    cfg.update(start_pos=[0, 0, 1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
    sim = pkg.QuadXSim(n, cfg, seed=1234, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.rand(K, n, 4, generator=g) * 2 - 1
    a[..., :3] *= 0.3
    a[..., 3] = (2 * HOVER_THR - 1) + 0.3 * a[..., 3]
    a = a.to(dev)
    obs = torch.zeros(K, n, 20, device=dev)
    rew = torch.zeros(K, n, device=dev)
    te = torch.zeros(K, n, dtype=torch.uint8, device=dev)
    tr = torch.zeros(K, n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sim.reset(obs[0])
    out = {"envs": n, "note": "L2 flushed (256 MB memset) before each timed block; 64 env steps per block"}

    def timed(fn, reps=5):
        best = []
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            best.append(s.elapsed_time(e))
        return statistics.median(best)

    def per_step():
        for k in range(K):
            sim.step(a[k], obs[k], rew[k], te[k], tr[k])

    per_step()
    ms = timed(per_step)
    out["per_step_launch"] = {"env_steps_per_s": n * K / (ms * 1e-3), "us_per_step": 1e3 * ms / K}
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        per_step()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            per_step()
    torch.cuda.synchronize()
    ms = timed(graph.replay)
    out["cuda_graph"] = {"env_steps_per_s": n * K / (ms * 1e-3), "us_per_step": 1e3 * ms / K}
    ms = timed(lambda: sim.step_k(a, obs, rew, te, tr))
    out["step_k64"] = {"env_steps_per_s": n * K / (ms * 1e-3), "us_per_step": 1e3 * ms / K, "alg_bytes_per_env_step": 102}
    sim.close()
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
