"""Per-phase clock stamps of the policy-forward kernel (CTA 0, slot 0): where does a tile's time go?"""
import ctypes as C
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import __graft_entry__ as ge

ge.build()
from fpv_drone_rl_agent_b200 import _lib, ppo

L = C.CDLL(_lib.LIB_PATH)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = ppo.ActorCritic().cuda()
pol = ppo.PackedPolicy(m, "cuda")
obs = torch.randn(n, 20, device="cuda")
acts = torch.zeros(n, 4, device="cuda"); vals = torch.zeros(n, device="cuda"); lp = torch.zeros(n, device="cuda"); on = torch.zeros(n, 20, device="cuda")
buf = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
ppo.policy_forward(pol, obs, actions=acts, values=vals, log_probs=lp, obs_norm=on)
L.ppo_debug_phase_clock(C.c_void_p(buf.data_ptr()))
ppo.policy_forward(pol, obs, actions=acts, values=vals, log_probs=lp, obs_norm=on)
torch.cuda.synchronize()
L.ppo_debug_phase_clock(None)
b = buf.cpu().view(8, 16)
names = ["wait G_a", "finish(prev)+epi 1p", "L2p mma", "epi 2p", "L3p+L1v mma", "stage X(next)+epi 1v", "L2v mma", "epi 2v", "issue L3v+L1p"]
flat = buf.cpu()
t0, t1 = int(flat[127]), int(flat[126])  # zero unless the library was built with QX_NVCC_EXTRA=-DPPO_K2_TRACE
if t0 and int(flat[125]):
    names_p = ["weights staged", "biases + TMEM alloc + barrier init", "__syncthreads", "per-slot setup", "first X tile loaded + staged"]
    ts = [t0] + [int(flat[k]) for k in (125, 124, 123, 122, 121)]
    print("CTA 0 prologue (cycles):", ", ".join(f"{nm} {ts[i + 1] - ts[i]}" for i, nm in enumerate(names_p)))
if t0:
    print(f"CTA 0: kernel entry -> first tile stamp {int(b[0, 0]) - t0} cycles; entry -> all slots done {t1 - t0} cycles "
          f"({(t1 - t0) / 1965:.1f} us at 1965 MHz)")
n_it = max(1, min(6, -(-n // 128) // (148 * 3) + 1))
for it in range(0, n_it):
    d = [int(b[it, k + 1] - b[it, k]) for k in range(9)]
    print(f"tile iter {it}: total {int(b[it, 9] - b[it, 0])} cycles |", ", ".join(f"{nm} {v}" for nm, v in zip(names, d)))
