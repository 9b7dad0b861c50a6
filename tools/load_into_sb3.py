#!/usr/bin/env python
"""Reference-side loader: turn an archive written by fpv_drone_rl_agent_b200 (PPOTrainer.save_sb3 / ppo.export_sb3_zip) into
the two files the reference's playback script loads (test_hover.py:8-11):

    python tools/load_into_sb3.py hover_sb3.zip out_prefix        # -> out_prefix.zip (PPO.load), out_prefix.pkl (VecNormalize.load)

Run it where the reference's own environment is installed (stable-baselines3 2.7.0, gymnasium, PyFlyt: simulation/uv.lock);
none of them exist in the build container, so this file is NOT exercised by the test-suite -- what the tests do check is
that the archive holds SB3's parameter names / shapes and that an SB3-style MlpPolicy forward over them reproduces the
device policy (tests/test_gpu_update.py::test_sb3_archive_plays_back).  It uses only public SB3 API:
PPO(...).policy.load_state_dict, VecNormalize attributes obs_rms / ret_rms, model.save, VecNormalize.save.
"""
import io
import json
import sys
import zipfile

import numpy as np
import torch


def main(src: str, out_prefix: str) -> None:
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env import DummyVecEnv, VecNormalize

    sys.path.insert(0, "simulation")
    from hover import QuadXHoverEnv  # the reference env (hover.py:10)

    with zipfile.ZipFile(src) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        data = json.loads(z.read("data").decode())
        vn = dict(np.load(io.BytesIO(z.read("vecnormalize.npz"))))
    venv = VecNormalize(DummyVecEnv([lambda: QuadXHoverEnv()]), norm_obs=bool(vn["norm_obs"]), norm_reward=bool(vn["norm_reward"]),
                        clip_obs=float(vn["clip_obs"]), clip_reward=float(vn["clip_reward"]), gamma=float(vn["gamma"]), epsilon=float(vn["epsilon"]))
    for name in ("obs_rms", "ret_rms"):
        rms = getattr(venv, name)
        rms.mean = vn[f"{name}.mean"].astype(np.float64).reshape(rms.mean.shape)
        rms.var = vn[f"{name}.var"].astype(np.float64).reshape(rms.var.shape)
        rms.count = float(vn[f"{name}.count"])
    model = PPO("MlpPolicy", venv, policy_kwargs=dict(net_arch=data["policy_kwargs"]["net_arch"], log_std_init=data["policy_kwargs"]["log_std_init"]),
                learning_rate=data["learning_rate"], n_steps=data["n_steps"], batch_size=min(data["batch_size"], data["n_steps"]),
                gamma=data["gamma"], gae_lambda=data["gae_lambda"], clip_range=data["clip_range"], ent_coef=data["ent_coef"],
                vf_coef=data["vf_coef"], max_grad_norm=data["max_grad_norm"], device="cpu")
    missing = model.policy.load_state_dict(sd, strict=True)
    print("policy.load_state_dict:", missing)
    model.num_timesteps = int(data["num_timesteps"])
    model.save(out_prefix)        # out_prefix.zip, what SAC/PPO.load of test_hover.py:11 opens
    venv.save(out_prefix + ".pkl")  # what VecNormalize.load of test_hover.py:8 opens
    print("wrote", out_prefix + ".zip", out_prefix + ".pkl")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
