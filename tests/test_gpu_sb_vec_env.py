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
    # the Box bounds themselves, not only the shapes: docking_env.py:84-95
    lo = np.array([-np.inf, -np.inf, -np.inf, -100, -100, -100, -np.pi, -np.pi / 2, -np.pi, -10 * np.pi, -10 * np.pi, -10 * np.pi], np.float32)
    assert np.array_equal(env.observation_space.low, lo) and np.array_equal(env.observation_space.high, -lo)
    assert env.observation_space.dtype == np.float32 and env.action_space.dtype == np.float32
    assert np.array_equal(env.action_space.low, -np.ones(4, np.float32)) and np.array_equal(env.action_space.high, np.ones(4, np.float32))
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
                # like the reference's, the info dict of a finished env describes the TERMINAL step (docking_env.py:226-229):
                # the pre-reset 13-states come from the kernel's terminal_chaser / terminal_target rows
                assert np.max(np.abs(info["chaser"] - g["chaser"][k, t])) < 1e-9 and np.max(np.abs(info["target"] - g["target"][k, t])) < 1e-9
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


def test_hover_spaces_match_reference():
    """hovering-v0: observation Box = Drone.state_lim_low/high (dynamics/quadrotor.py:35-39 via hovering_env.py:36-40), action
    Box = [0, 1]^4 (hovering_env.py:42)."""
    env = SBVecEnv("hovering-v0", 4)
    lo = np.array([-100, -100, 0, -100, -100, -100, -100, -100, -100, -100, -10 * 2 * np.pi, -10 * 2 * np.pi, -10 * 2 * np.pi], np.float32)
    hi = np.array([100] * 10 + [10 * 2 * np.pi] * 3, np.float32)
    assert np.array_equal(env.observation_space.low, lo) and np.array_equal(env.observation_space.high, hi)
    assert np.array_equal(env.action_space.low, np.zeros(4, np.float32)) and np.array_equal(env.action_space.high, np.ones(4, np.float32))
    env.close()
