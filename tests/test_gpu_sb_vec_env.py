"""SBVecEnv: the stable-baselines-shaped numpy face (what SubprocVecEnv hands to PPO2 in the reference,
run_moving_docking_ppo2.py:65-67, rl_baselines/ppo2/ppo2.py:472-498) replays the reference's random-action tape: float32
observations, numpy rewards / dones, a list of info dicts with the reference's keys (docking_env.py:226-229) and SB's
terminal_observation, auto-reset included."""
import numpy as np
import pytest
import torch

from qs_testutil import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from quadsim_b200.sb_vec_env import SBVecEnv


@pytest.mark.parametrize("name,env_id", [("dock_v2_rand", "docking-v2"), ("dock_v0_rand", "docking-v0")])
def test_sb_vec_env_replays_reference_tape(name, env_id):
    g = load_golden(name)
    K, T = g["actions"].shape[:2]
    env = SBVecEnv(env_id, K, seed=int(g["seed"]), env_id0=int(g["env_id0"]), randomize="episode", half_width=float(g["half_width"]))
    assert env.num_envs == K and env.observation_space.shape == (12,) and env.action_space.shape == (4,)
    obs = env.reset()
    assert obs.dtype == np.float32 and obs.shape == (K, 12) and np.allclose(obs, g["obs0"], atol=1e-6)
    n_done = 0
    for t in range(300):
        obs, rew, done, infos = env.step(g["actions"][:, t])
        assert obs.dtype == np.float32 and rew.shape == (K,) and done.dtype == np.bool_ and isinstance(infos, list) and len(infos) == K
        assert np.array_equal(done, g["done"][:, t] != 0), t
        assert np.allclose(obs, g["obs"][:, t], atol=2e-6, rtol=1e-6) and np.allclose(rew, g["reward"][:, t], atol=1e-6)
        for k in range(K):
            info = infos[k]
            assert set(info) >= {"chaser", "target", "flag_docking", "done_overlimit"}
            assert info["flag_docking"] == bool(g["docking"][k, t]) and info["done_overlimit"] == bool(g["overlimit"][k, t])
            if done[k]:
                n_done += 1
                assert np.allclose(info["terminal_observation"], g["term_obs"][k, t], atol=2e-6, rtol=1e-6)
                assert np.array_equal(info["chaser"][:3], g["ini_pos"][k, t])  # re-seated (documented in sb_vec_env.py)
            else:
                assert np.max(np.abs(info["chaser"] - g["chaser"][k, t])) < 1e-9 and np.max(np.abs(info["target"] - g["target"][k, t])) < 1e-9
    assert n_done >= 4
    env.close()


def test_sb_vec_env_monitor_episode_info():
    """monitor=True: info['episode'] = {'r', 'l'} on done (SB's Monitor; the key rl_baselines/ppo2/ppo2.py:494-497 reads) equals the
    reference tape's per-episode reward sum and length."""
    g = load_golden("dock_v2_rand")
    K = g["actions"].shape[0]
    env = SBVecEnv("docking-v2", K, seed=int(g["seed"]), env_id0=int(g["env_id0"]), randomize="episode", half_width=float(g["half_width"]),
                   monitor=True)
    env.reset()
    ret, length, seen = np.zeros(K), np.zeros(K, int), 0
    for t in range(300):
        _, rew, done, infos = env.step(g["actions"][:, t])
        ret += g["reward"][:, t]
        length += 1
        for k in range(K):
            assert ("episode" in infos[k]) == bool(done[k])
            if done[k]:
                assert infos[k]["episode"]["l"] == length[k] and abs(infos[k]["episode"]["r"] - ret[k]) < 1e-6
                ret[k], length[k], seen = 0.0, 0, seen + 1
    assert seen >= 4
    env.close()
