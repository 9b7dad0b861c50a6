"""One eager (no CUDA graph) rollout of a few steps, for an ncu launch list of the per-step kernels:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rollout_launches.csv \
        python tools/rollout_launch_list.py [envs] [steps]
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import __graft_entry__ as ge

ge.build()
from fpv_drone_rl_agent_b200 import ppo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
t = ppo.PPOTrainer(ppo.PPOConfig(n_envs=n, n_steps=T, seed=0, use_cuda_graph=False), device="cuda")
for _ in range(3):
    t.rollout.collect()
torch.cuda.synchronize()
print("ok")
