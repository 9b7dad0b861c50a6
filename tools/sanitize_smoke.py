"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
    gpurun -- 'timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py'"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import __graft_entry__ as ge

ge.build()
import fpv_drone_rl_agent_b200 as pkg
from fpv_drone_rl_agent_b200 import ppo

n = 257
env = pkg.QuadXHoverVecEnv(n, seed=1, infos=False)
env.reset()
a = torch.zeros(n, 4, device="cuda"); a[:, 3] = -1.0
for k in range(34):  # crosses the mass termination on step 32 (reset queue, warp-aggregated atomics)
    env.step(a)
m = torch.zeros(n, dtype=torch.uint8, device="cuda"); m[::3] = 1
env.sim.reset(env.obs, m)
K = 3
acts = torch.rand(K, n, 4, device="cuda") - 0.5
obs = torch.zeros(K, n, 20, device="cuda"); rew = torch.zeros(K, n, device="cuda")
te = torch.zeros(K, n, dtype=torch.uint8, device="cuda"); tr = torch.zeros_like(te)
env.sim.step_k(acts, obs, rew, te, tr)
o, r, t1, t2, _ = env.sim.step_host(np.zeros((n, 4), np.float32), want_terminal_obs=True)
env.close()
y = pkg.QuadXYawVecEnv(130, seed=2)
y.reset()
for k in range(5):
    y.step(torch.rand(130, 1, device="cuda") - 0.5)
y.close()
cfg = ppo.PPOConfig(n_envs=300, n_steps=20, seed=0, n_epochs=1, batch_size=2048, use_cuda_graph=False)
t = ppo.PPOTrainer(cfg, device="cuda")
print(t.learn_iteration())
T, nn = 40, 77
adv = torch.zeros(T, nn, device="cuda"); ret = torch.zeros(T, nn, device="cuda")
ppo.gae(torch.randn(T, nn, device="cuda"), torch.randn(T, nn, device="cuda"), torch.zeros(T, nn, dtype=torch.uint8, device="cuda"),
        torch.randn(nn, device="cuda"), 0.99, 0.95, adv, ret)
torch.cuda.synchronize()
print("sanitize smoke done")
