"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (gradient all-reduce, merge of the
normalisation statistics, rank handling of bench.py's reference arm).  The data path itself has no collective."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fpv_drone_rl_agent_b200 import ppo

    # (1) replicas start identical, see different data, and stay identical after all-reduced steps
    torch.manual_seed(0)
    model = ppo.ActorCritic()
    opt = torch.optim.Adam(model.parameters(), lr=3e-4, eps=1e-5)
    g = torch.Generator().manual_seed(100 + rank)
    local_grads = None
    for it in range(3):
        obs, act = torch.randn(256, 20, generator=g), torch.randn(256, 4, generator=g)
        v, logp, _ = model.evaluate_actions(obs, act)
        loss = (v**2).mean() - logp.mean()
        opt.zero_grad()
        loss.backward()
        if it == 0:
            local_grads = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        ppo.allreduce_gradients(model.parameters(), world)
        if it == 0:
            avg = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    lg = [torch.zeros_like(local_grads) for _ in range(world)]
    dist.all_gather(lg, local_grads)
    ok_rep = all(torch.equal(gathered[0], x) for x in gathered)
    ok_avg = torch.allclose(avg, sum(lg) / world, atol=1e-7)
    # (2) merge of per-rank running statistics == statistics of the pooled data
    rs = ppo.RunningStats(3, "cpu")
    x = torch.randn(500 + 100 * rank, 3, generator=g, dtype=torch.float64) * (1 + rank) + rank
    rs.stats[:3], rs.stats[3:6], rs.stats[6] = x.mean(0), x.var(0, unbiased=False), float(x.shape[0])
    xs = [torch.zeros(600, 3, dtype=torch.float64) for _ in range(world)]
    pad = torch.zeros(600, 3, dtype=torch.float64); pad[: x.shape[0]] = x
    dist.all_gather(xs, pad)
    pooled = torch.cat([xs[r][: 500 + 100 * r] for r in range(world)])
    rs.allreduce_()
    ok_stats = torch.allclose(rs.stats[:3], pooled.mean(0), atol=1e-12) and torch.allclose(rs.stats[3:6], pooled.var(0, unbiased=False), atol=1e-12)
    ok_ids = ppo.shard_env_ids(rank, 4096) == rank * 4096
    q.put((rank, ok_rep, ok_avg, ok_stats, ok_ids, flat.numel()))
    dist.destroy_process_group()


def test_gradient_allreduce_and_stats_merge_world2():
    mp.set_start_method("spawn", force=True)
    q = mp.get_context("spawn").Queue()
    port = 29500 + os.getpid() % 1000
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    for rank, ok_rep, ok_avg, ok_stats, ok_ids, nparams in res:
        assert ok_rep and ok_avg and ok_stats and ok_ids, (rank, ok_rep, ok_avg, ok_stats, ok_ids)
        assert nparams == 39049  # actor 19 716 + log_std 4 + critic 19 329


def test_reference_arm_rank_handling():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    outs = []
    for rank in (0, 1):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
                            "--cpu-envs", "256"], env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    line = json.loads(outs[0])
    assert outs[1] == ""
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["value"] > 0 and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_narrow_net_arch_is_an_exact_embedding():
    """net_arch=[64, 64] (SB3's default) is held zero-padded in the 128-wide layers the kernels are tiled for: the
    forward pass and the gradients are those of a genuine 64-wide network, the padded entries get exactly zero gradient, and the
    SB3 export has the narrow shapes."""
    sys.path.insert(0, ROOT)
    from fpv_drone_rl_agent_b200 import ppo

    torch.manual_seed(11)
    m = ppo.ActorCritic(12, 1, log_std_init=-0.5, hidden=64)
    sd = ppo.export_sb3_state_dict(m)
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (64, 12) and sd["mlp_extractor.policy_net.2.weight"].shape == (64, 64)
    assert sd["action_net.weight"].shape == (1, 64) and sd["value_net.weight"].shape == (1, 64)
    assert sum(v.numel() for v in sd.values()) == m.num_effective_params() == 2 * (12 * 64 + 64 + 64 * 64 + 64) + 64 + 1 + 64 + 1 + 1
    # a genuine 64-wide torch network with the exported weights
    pi = torch.nn.Sequential(torch.nn.Linear(12, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1))
    vf = torch.nn.Sequential(torch.nn.Linear(12, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1))
    with torch.no_grad():
        for net, names in ((pi, ("mlp_extractor.policy_net.0", "mlp_extractor.policy_net.2", "action_net")),
                           (vf, ("mlp_extractor.value_net.0", "mlp_extractor.value_net.2", "value_net"))):
            for k, name in zip((0, 2, 4), names):
                net[k].weight.copy_(sd[name + ".weight"]); net[k].bias.copy_(sd[name + ".bias"])
    x, a = torch.randn(64, 12), torch.randn(64, 1)
    mean, value = m(x)
    assert torch.allclose(mean, pi(x), atol=1e-6) and torch.allclose(value, vf(x).squeeze(-1), atol=1e-6)
    v, logp, _ = m.evaluate_actions(x, a)
    (logp.mean() + (v ** 2).mean()).backward()
    z = (a - pi(x)) / m.log_std.detach().exp()
    ((-0.5 * z * z - m.log_std.detach() - 0.5 * np.log(2 * np.pi)).sum(-1).mean() + (vf(x).squeeze(-1) ** 2).mean()).backward()
    assert torch.allclose(m.pi2.weight.grad[:64, :64], pi[2].weight.grad, atol=1e-6) and torch.allclose(m.vf1.weight.grad[:64], vf[0].weight.grad, atol=1e-6)
    for lin in (m.pi1, m.pi2, m.vf1, m.vf2):
        assert float(lin.weight.grad[64:].abs().max()) == 0.0 and float(lin.bias.grad[64:].abs().max()) == 0.0
    for lin in (m.pi2, m.vf2, m.mu, m.v):
        assert float(lin.weight.grad[:, 64:].abs().max()) == 0.0
    m2 = ppo.ActorCritic(12, 1, hidden=64)
    ppo.import_sb3_state_dict(m2, sd)
    assert torch.equal(m2(x)[0], mean.detach()) and m2.pi2.weight[64:].abs().max() == 0


def test_sb3_compatible_export_roundtrip():
    """train_hover.py:26-27,62-63 save an SB3 zip + VecNormalize pkl; the export uses SB3's parameter names."""
    sys.path.insert(0, ROOT)
    from fpv_drone_rl_agent_b200 import ppo

    torch.manual_seed(3)
    m = ppo.ActorCritic()
    sd = ppo.export_sb3_state_dict(m)
    assert set(sd) == {"log_std", "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
                       "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                       "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "action_net.weight", "action_net.bias",
                       "value_net.weight", "value_net.bias"}
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (128, 20) and sd["action_net.weight"].shape == (4, 128) and sd["value_net.weight"].shape == (1, 128)
    assert sum(v.numel() for v in sd.values()) == 39049
    m2 = ppo.ActorCritic()
    ppo.import_sb3_state_dict(m2, sd)
    x = torch.randn(5, 20)
    assert torch.equal(m(x)[0], m2(x)[0]) and torch.equal(m(x)[1], m2(x)[1])
    rs, rr = ppo.RunningStats(20, "cpu"), ppo.RunningStats(1, "cpu")
    vn = ppo.export_vecnormalize(rs, rr, ppo.PPOConfig())
    assert vn["obs_rms"]["mean"].shape == (20,) and vn["obs_rms"]["count"] == 1e-4 and vn["clip_obs"] == 10.0 and vn["gamma"] == 0.99


def test_sb3_style_archive(tmp_path):
    """export_sb3_zip lays the policy out like SB3's save_to_zip_file (policy.pth + version file), plus the VecNormalize
    statistics as arrays; what comes back out of the archive is the same network."""
    import io
    import zipfile

    import numpy as np

    sys.path.insert(0, ROOT)
    from fpv_drone_rl_agent_b200 import ppo

    torch.manual_seed(5)
    m = ppo.ActorCritic(log_std_init=-1.0)
    rs, rr = ppo.RunningStats(20, "cpu"), ppo.RunningStats(1, "cpu")
    path = str(tmp_path / "hover_b200.zip")
    ppo.export_sb3_zip(m, path, ppo.export_vecnormalize(rs, rr, ppo.PPOConfig()))
    with zipfile.ZipFile(path) as z:
        assert {"policy.pth", "_stable_baselines3_version", "vecnormalize.npz"} <= set(z.namelist())
        assert z.read("_stable_baselines3_version").decode() == "2.7.0"  # the reference's pin, uv.lock
        sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True)
        vn = np.load(io.BytesIO(z.read("vecnormalize.npz")))
    assert sum(v.numel() for v in sd.values()) == 39049 and float(sd["log_std"][0]) == -1.0
    m2 = ppo.ActorCritic()
    ppo.import_sb3_state_dict(m2, sd)
    x = torch.randn(7, 20)
    assert torch.equal(m(x)[0], m2(x)[0]) and torch.equal(m(x)[1], m2(x)[1])
    assert vn["obs_rms.mean"].shape == (20,) and vn["ret_rms.var"].shape == (1,) and float(vn["clip_obs"]) == 10.0
