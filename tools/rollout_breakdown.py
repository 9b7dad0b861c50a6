"""Where does a rollout step go?  Times RolloutEngine.collect with pieces switched off (131 072 envs x 32 steps)."""
import statistics, sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import __graft_entry__ as ge; ge.build()
from fpv_drone_rl_agent_b200 import ppo
n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 32
def run(**kw):
    cfg = ppo.PPOConfig(n_envs=n, n_steps=T, seed=0, **kw)
    t = ppo.PPOTrainer(cfg, device="cuda")
    for _ in range(3): t.rollout.collect()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); t.rollout.collect(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    t.sim.close()
    return statistics.median(ts) / T * 1e3
full = run()
print(f"full rollout step: {full:.1f} us")
print(f"  without obs normalisation (2 launches): {run(norm_obs=False):.1f} us")
print(f"  without reward normalisation (4 launches -> 2 torch ops): {run(norm_reward=False):.1f} us")
print(f"  without both: {run(norm_obs=False, norm_reward=False):.1f} us")
