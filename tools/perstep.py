import sys, json, torch, numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge; ge.build()
import fpv_drone_rl_agent_b200 as pkg
HOVER_THR = (0.1*9.81/4.0)**0.5
E = 1<<20
cfg = pkg.default_config(); cfg.update(start_pos=[0,0,1.0], spawn_throttle=HOVER_THR, auto_reset=1, noise=1)
sim = pkg.QuadXSim(E, cfg, seed=1234)
dev = sim.device
g = torch.Generator(device="cpu").manual_seed(0)
acts = torch.rand(8, E, 4, generator=g)*2-1; acts[..., :3] *= 0.3; acts[..., 3] = (2*HOVER_THR-1) + 0.3*acts[..., 3]
acts = acts.to(dev)
obs = torch.zeros(E,20,device=dev); rew=torch.zeros(E,device=dev); te=torch.zeros(E,dtype=torch.uint8,device=dev); tr=torch.zeros(E,dtype=torch.uint8,device=dev)
sim.reset(obs)
K=120
ev=[torch.cuda.Event(enable_timing=True) for _ in range(K+1)]
dones=[]
ev[0].record()
for k in range(K):
    sim.step(acts[k%8], obs, rew, te, tr); ev[k+1].record()
    dones.append((te|tr).sum())
torch.cuda.synchronize()
per=[ev[k].elapsed_time(ev[k+1]) for k in range(K)]
for k in range(0,K,4): print(k, ["%.3f/%.1f%%"%(per[j], 100*float(dones[j])/E) for j in range(k,k+4)])
