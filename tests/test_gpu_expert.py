"""GPU checks of the batched PID expert (qs_expert_action) through the C ABI: the reference's own expert loop
(tests/golden/expert_v0.npz: two episodes incl. the scripts' state_des = env.chaser_ini_state aliasing), the fp64 oracle at
larger N on docking-v1 / docking-v2, and the recorded-dataset layout of generate_PID_expert_traj (run_expert_record.py:177-183).
fp64: 1e-9 and exact flags; fp32: 1e-4 * max(1, |ref|)."""
import numpy as np
import pytest
import torch

from qs_testutil import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from quadsim_b200.expert import PIDExpert, generate_pid_expert_traj
    from oracle import oracle as O


def rel_err(a, ref):
    a = np.asarray(a, dtype=np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(1.0, np.abs(ref))))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
def test_expert_replays_reference_tape(dtype, tol):
    g = load_golden("expert_v0")
    env = qs.make_vec("docking-v0", 1, dtype=dtype)
    ex = PIDExpert(env, alias_ini_state=True)
    obs = env.reset()
    for t in range(len(g["actions"])):
        act = ex.act(obs)
        assert rel_err(obs.cpu().numpy()[0], g["obs"][t]) <= tol and rel_err(act.cpu().numpy()[0], g["actions"][t]) <= tol, t
        obs, rew, done, info = env.step(act)
        assert bool(done[0]) == bool(g["done"][t]) and bool(info.flag_docking[0]) == bool(g["docking"][t]), t
        assert rel_err(rew.cpu().numpy()[0], g["rewards"][t]) <= tol, t
        if t % 50 == 0 and not done[0]:
            st = env.get_state()
            assert rel_err(st["chaser"].cpu().numpy()[0], g["chaser"][t]) <= tol and rel_err(st["target"].cpu().numpy()[0], g["target"][t]) <= tol, t
    s = env.stats()
    assert s["episodes"] == 2 and s["docked_steps"] == float(g["docking"].sum())
    assert abs(s["return_sum"] - g["episode_returns"].sum()) <= 600 * tol
    env.close()


@pytest.mark.parametrize("env_id,alias", [("docking-v1", True), ("docking-v1", False), ("docking-v2", True)])
def test_expert_vs_oracle_batched(env_id, alias):
    n, T = 300, 700  # ragged (2.3 tiles); crosses every env's first reset
    kind, rnd, hw = {"docking-v1": ("docking-v0", O.RAND_ONCE, 0.3), "docking-v2": ("docking-v2", O.RAND_ONCE, 0.5)}[env_id]
    env = qs.make_vec(env_id, n, dtype=torch.float64, seed=21, randomize="once", half_width=hw)
    ref = O.OracleVecEnv(kind, n, seed=21, randomize=rnd, half_width=hw)
    ex, rex = PIDExpert(env, alias), O.OraclePIDExpert(ref, alias)
    obs, robs = env.reset(), ref.reset()
    for t in range(T):
        act, ract = ex.act(obs), rex.act(robs)
        assert rel_err(obs.cpu().numpy(), robs) <= 1e-9 and rel_err(act.cpu().numpy(), ract) <= 1e-9, t
        obs, rew, done, info = env.step(act)
        robs, rrew, rdone, rinfo = ref.step(ract)
        assert np.array_equal(done.cpu().numpy(), rdone != 0) and rel_err(rew.cpu().numpy(), rrew) <= 1e-9, t
    assert rel_err(ex.mem.cpu().numpy(), rex.mem) <= 1e-9
    assert env.stats()["episodes"] >= n  # docking-v0/v1 flights end by overtime at 600; v2's chase may leave the 10 m ball
    env.close()
    ref.close()


def test_generate_pid_expert_traj_layout(tmp_path):
    n = 8
    env = qs.make_vec("docking-v1", n, dtype=torch.float64, seed=3)
    data = generate_pid_expert_traj(env, n_episodes=16, save_path=str(tmp_path / "expert_pid"))
    assert set(data) == {"actions", "obs", "rewards", "episode_returns", "episode_starts"}  # run_expert_record.py:177-183
    m = len(data["actions"])
    assert data["obs"].shape == (m, 12) and data["actions"].shape == (m, 4) and data["rewards"].shape == (m,)
    assert data["episode_starts"].sum() == 16 and data["episode_starts"][0] and data["episode_returns"].shape == (16,)
    bounds = np.r_[np.nonzero(data["episode_starts"])[0], m]
    assert np.all(np.diff(bounds) <= 600) and np.any(np.diff(bounds) == 600)  # overtime at 600 (docking_env.py:152) or earlier
    for a, b, r in zip(bounds[:-1], bounds[1:], data["episode_returns"]):
        assert abs(data["rewards"][a:b].sum() - r) < 1e-9
    assert np.allclose(data["obs"][0, 3:], 0) and abs(data["obs"][0, 0] - 1.8) < 0.31  # env 0's first reset observation
    assert np.median(data["episode_returns"]) > 0.5  # the expert docks: the reference's own flight returns 0.84
    loaded = np.load(str(tmp_path / "expert_pid.npz"))
    assert all(np.array_equal(loaded[k], data[k]) for k in data)
    env.close()
