#!/usr/bin/env python
"""bench.py's bounded PPO training-to-target leg, repeated: the spread of time-to-target over runs of the same seed
(the update's floating-point atomics make the trajectories differ from run to run).  Usage: python tools/train_leg_repeat.py [runs] [budget_s]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import bench  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
for k in range(runs):
    t = bench.train_leg(torch.device("cuda", 0), 0, 1, budget_s=budget)
    print(json.dumps({k2: t[k2] for k2 in ("iterations", "wall_s", "time_to_ep_len_s", "time_to_ep_len_and_rew_s", "final_ep_rew_mean", "best_ep_rew_mean")}), flush=True)
