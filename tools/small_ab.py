#!/usr/bin/env python
"""bench.py's configs[1] legs (4096 envs: per-step launches, CUDA graph, step_k) for the current QX_* settings."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import fpv_drone_rl_agent_b200 as pkg  # noqa: E402
import bench  # noqa: E402

r = bench.small_config(pkg, torch.device("cuda", 0))
print(json.dumps({"QX_PDL": os.environ.get("QX_PDL", "1"), "launch_us": round(r["per_step_launch"]["us_per_step"], 2),
                  "graph_us": round(r["cuda_graph"]["us_per_step"], 2), "step_k_us": round(r["step_k64"]["us_per_step"], 2)}))
