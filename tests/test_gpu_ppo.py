"""GPU checks of the device-resident PPO2 pieces around the CUDA env (BASELINE config 5): the GAE kernel
(bit-exact against the reference loop and the oracle), the shipped policy flying batched docking-v0 into the docked regime,
rollout bookkeeping, and a short learn() run."""
import numpy as np
import pytest
import torch

from qs_testutil import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from quadsim_b200 import ppo
    from oracle import oracle as O


def test_gae_kernel_bit_exact():
    g = load_golden("gae_numpy")
    dev = "cuda"
    adv, ret = ppo.gae(torch.as_tensor(g["rew"], device=dev), torch.as_tensor(g["val"], device=dev),
                       torch.as_tensor(g["dones"], device=dev), torch.as_tensor(g["last_val"], device=dev),
                       torch.as_tensor(g["last_d"], device=dev), float(g["gamma"]), float(g["lam"]))
    assert np.array_equal(adv.cpu().numpy(), g["adv"]) and np.array_equal(ret.cpu().numpy(), g["ret"])
    # larger, ragged, vs the oracle
    rng = np.random.RandomState(0)
    T, n = 96, 10007
    rew, val = rng.normal(size=(T, n)).astype(np.float32), rng.normal(size=(T, n)).astype(np.float32)
    dones, lv, ld = rng.rand(T, n) < 0.02, rng.normal(size=n).astype(np.float32), rng.rand(n) < 0.02
    a_ref, r_ref = O.gae(rew, val, dones, lv, ld, 0.99, 0.95)
    adv, ret = ppo.gae(*[torch.as_tensor(x, device=dev) for x in (rew, val, dones, lv, ld)], 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), a_ref) and np.array_equal(ret.cpu().numpy(), r_ref)


def _shipped_policy():
    g = load_golden("policy_v0")
    return ppo.MlpPolicy().load_sb_params({k: g["param::" + k] for k in g["param_order"]}).cuda()


def test_shipped_policy_docks_batched_env():
    """The reference's trained policy (trained_model/best_model_v0.zip) on the CUDA docking-v0: the reference run docks on
    183 of 600 steps and ends by overtime with return ~0.71.  GPU matmul rounding differs from numpy's, so
    the closed loop is compared statistically, and a jittered batch must stay in the same regime."""
    pol = _shipped_policy()
    n = 512
    env = qs.make_vec("docking-v0", n, dtype=torch.float64, seed=4, randomize="episode", half_width=0.05)
    obs = env.reset()
    docked = torch.zeros(n, device="cuda")
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(600):
        obs, rew, done, info = env.step(pol.predict(obs.float()).double())
        docked += info.flag_docking
        ret += rew
        if t < 599:
            assert not bool(done.any()), t
    assert bool(done.all()) and bool(info.done_overtime.all())
    # measured: mean 146 docked steps/episode, 84 % of the +-5 cm starts dock for > 100 steps (the policy was trained from one start)
    assert 110 <= float(docked.mean()) <= 200 and float((docked > 100).float().mean()) > 0.7, float(docked.mean())
    assert float(ret.max()) < 1.5 and float(ret.mean()) > -1.0
    s = env.stats()
    assert s["episodes"] == n and s["mean_length"] == 600 and s["docked_steps"] == float(docked.sum())
    assert abs(s["mean_return"] - float(ret.mean())) < 1e-9
    env.close()


def test_runner_bookkeeping_and_gae_vs_oracle():
    n, T = 1024, 48
    env = qs.make_vec("docking-v2", n, seed=2, randomize="episode")
    torch.manual_seed(0)
    pol = ppo.MlpPolicy().cuda()
    with torch.no_grad():
        pol.logstd.fill_(0.5)  # noisy actions -> some episodes end inside the rollout
    runner = ppo.DeviceRunner(env, pol, T)
    b1 = runner.run()
    b2 = runner.run()  # continues from where the first rollout stopped (stock SB; the in-tree fork resets instead)
    assert b2["obs"].shape == (T * n, 12) and b2["actions"].shape == (T * n, 4)
    assert torch.equal(runner.mb_dones[0], (b1["rewards"][-1] * 0).to(torch.uint8) + runner.mb_dones[0])  # uint8 flags
    a_ref, r_ref = O.gae(runner.mb_rewards.cpu().numpy(), runner.mb_values.cpu().numpy(), runner.mb_dones.cpu().numpy(),
                         pol.forward(runner.obs)[1].detach().cpu().numpy(), runner.dones.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(b2["advs"].view(T, n).cpu().numpy(), a_ref)
    assert np.array_equal(b2["returns"].view(T, n).cpu().numpy(), r_ref)
    assert int(runner.mb_dones.sum()) > 0
    env.close()


def test_ppo2_learn_runs_and_updates():
    n = 2048
    env = qs.make_vec("docking-v2", n, seed=3, randomize="episode")
    algo = ppo.PPO2(env, n_steps=32, nminibatches=4, noptepochs=2, seed=1)
    before = [p.detach().clone() for p in algo.policy.parameters()]
    hist = algo.learn(3 * n * 32)
    assert len(hist) == 3
    for h in hist:
        assert all(np.isfinite(v) for k, v in h.items() if k not in ("ep_mean_return", "ep_mean_length")), h
        assert h["approxkl"] < 0.05 and 0 <= h["clipfrac"] <= 1
    assert any(not torch.equal(a, b) for a, b in zip(before, algo.policy.parameters()))
    assert hist[-1]["vf_loss"] < hist[0]["vf_loss"]  # the value head fits the (large, negative) shaped returns
    env.close()


@pytest.mark.parametrize("n_slices,precision", [(1, "bf16"), (2, "bf16"), (3, "bf16"), (8, "bf16"), (2, "fp32")])
def test_in_place_rollout_equals_stepwise_loop(n_slices, precision):
    """qs_rollout (policy + env kernels writing the rollout rows in place) against the same kernels driven step by step
    through the public objects on a twin environment: every stored quantity must be bit-identical, over two consecutive
    rollouts (the second continues from row T of the first).  Also with the fp32 forward (qs_policy_set_precision)."""
    from quadsim_b200.fused_policy import FusedPolicy
    n, T = 3000, 24  # ragged: 23.4 tiles
    torch.manual_seed(0)
    pol = ppo.MlpPolicy().cuda()
    with torch.no_grad():
        pol.logstd.fill_(0.5)  # noisy actions -> some episodes end (and auto-reset) inside the rollouts
        pol.pi.weight.mul_(20.0)
    env_a = qs.make_vec("docking-v2", n, seed=5, randomize="episode", want_info=False)
    env_b = qs.make_vec("docking-v2", n, seed=5, randomize="episode", want_info=False)
    fa, fb = FusedPolicy(pol, seed=11, precision=precision), FusedPolicy(pol, seed=11, precision=precision)
    runner = ppo.DeviceRunner(env_a, pol, T, fused=fa, n_slices=n_slices)
    assert runner.in_place
    obs = env_b.reset().float()
    assert torch.equal(obs, runner.obs)
    dones = torch.zeros(n, dtype=torch.uint8, device="cuda")
    n_done = 0
    for rollout in range(5):
        batch = runner.run()
        n_done += int(runner._dones_all[1:].sum())
        for t in range(T):
            act, clip, val, nlp = fb.step(obs)
            assert torch.equal(runner.mb_obs[t], obs) and torch.equal(runner.mb_dones[t], dones), (rollout, t)
            assert torch.equal(runner.mb_actions[t], act) and torch.equal(runner.mb_values[t], val) and torch.equal(runner.mb_neglogp[t], nlp), (rollout, t)
            obs, rew, done, _ = env_b.step(clip)
            dones = done.to(torch.uint8)
            assert torch.equal(runner.mb_rewards[t], rew), (rollout, t)
        assert torch.equal(runner.obs, obs) and torch.equal(runner.dones, dones)
        last_v = fb.step(obs, deterministic=True, step_idx=0)[2]
        a_ref, r_ref = O.gae(runner.mb_rewards.cpu().numpy(), runner.mb_values.cpu().numpy(), runner.mb_dones.cpu().numpy(),
                             last_v.cpu().numpy(), dones.cpu().numpy(), 0.99, 0.95)
        assert np.array_equal(batch["advs"].view(T, n).cpu().numpy(), a_ref) and np.array_equal(batch["returns"].view(T, n).cpu().numpy(), r_ref)
    assert n_done > 0
    env_a.close()
    env_b.close()


@pytest.mark.parametrize("env_id", ["docking-v2", "docking-v0"])
def test_fused_step_kernel_rollout_equals_two_kernel_rollout(env_id):
    """qs_rollout(n_slices=0) -- ONE launch per step: the policy half of every CTA hands the clipped actions to env warps of the
    same CTA, which run the env step kernel's own device code -- against qs_rollout(n_slices=1) (two launches per step) on a
    twin environment, five consecutive rollouts with auto-resets on a ragged env count.  The policy arithmetic is the same
    instruction sequence (bit-identical actions / values / log-probs for identical observations); the env arithmetic is the
    same source compiled into another kernel, so it is compared within 1e-5 relative, done flags exactly."""
    from quadsim_b200.fused_policy import FusedPolicy
    n, T = 128 * 9 + 77, 24
    torch.manual_seed(0)
    pol = ppo.MlpPolicy().cuda()
    with torch.no_grad():
        pol.logstd.fill_(0.5)
        pol.pi.weight.mul_(20.0)
    envs = [qs.make_vec(env_id, n, seed=5, randomize="episode", want_info=False) for _ in range(2)]
    fps = [FusedPolicy(pol, seed=11) for _ in range(2)]
    fused = ppo.DeviceRunner(envs[0], pol, T, fused=fps[0], n_slices=0)
    plain = ppo.DeviceRunner(envs[1], pol, T, fused=fps[1], n_slices=1)
    assert fused.in_place and plain.in_place and fused.launches_per_rollout() == T + 2 and plain.launches_per_rollout() == 2 * T + 2
    close_ = lambda x, y: bool(((x - y).abs() <= 1e-5 * y.abs().clamp(min=1)).all())
    identical, n_done = True, 0
    for rollout in range(5):
        ba, bb = fused.run(), plain.run()
        n_done += int(fused._dones_all[1:].sum())
        assert torch.equal(fused._dones_all, plain._dones_all), rollout
        for name in ("_obs_all", "mb_rewards", "mb_actions", "_values_all", "mb_neglogp"):
            x, y = getattr(fused, name), getattr(plain, name)
            assert close_(x, y), (rollout, name, float((x - y).abs().max()))
            identical &= torch.equal(x, y)
        assert close_(ba["advs"], bb["advs"]) and close_(ba["returns"], bb["returns"])
        # teacher-forced: the policy half alone, on the fused run's own observations, is bit-identical to qs_policy_step
        for t in (0, T // 2, T - 1):
            act, clip, val, nlp = fps[1].step(fused.mb_obs[t], step_idx=rollout * T + t)
            assert torch.equal(act, fused.mb_actions[t]) and torch.equal(val, fused.mb_values[t]) and torch.equal(nlp, fused.mb_neglogp[t])
    sa, sb = envs[0].get_state(), envs[1].get_state()
    assert torch.equal(sa["t"], sb["t"]) and torch.equal(sa["episode"], sb["episode"]) and close_(sa["chaser"], sb["chaser"])
    st_a, st_b = envs[0].stats(), envs[1].stats()
    assert st_a["episodes"] == st_b["episodes"] == n_done > 0 and st_a["length_sum"] == st_b["length_sum"]
    print(env_id, "fused step kernel vs two-kernel rollout: bit-identical =", identical, "episodes", n_done)
    with pytest.raises(qs.QuadSimError):  # the fused kernel is built for the compiled-in model constants
        ppo.DeviceRunner(qs.make_vec(env_id, 256, params={"mass": 0.2}), pol, 4, fused=fps[0], n_slices=0).run()
    for e in envs:
        e.close()


def test_fused_step_kernel_refuses_fp32_policy():
    from quadsim_b200 import _capi
    from quadsim_b200.fused_policy import FusedPolicy
    pol = ppo.MlpPolicy().cuda()
    env = qs.make_vec("docking-v2", 512, seed=1, want_info=False)
    runner = ppo.DeviceRunner(env, pol, 4, fused=FusedPolicy(pol, precision="fp32"), n_slices=0)
    with pytest.raises(_capi.QuadSimError, match="QS_POLICY_BF16"):
        runner.run()
    env.close()


def test_ppo2_learn_with_fused_policy():
    n = 4096
    env = qs.make_vec("docking-v2", n, seed=3, randomize="episode")
    algo = ppo.PPO2(env, n_steps=32, nminibatches=4, noptepochs=2, seed=1, fused_policy=True)
    assert algo.runner.in_place
    hist = algo.learn(3 * n * 32)
    assert len(hist) == 3 and all(np.isfinite(h["pg_loss"]) and np.isfinite(h["vf_loss"]) for h in hist)
    assert all(h["approxkl"] < 0.05 for h in hist) and hist[-1]["vf_loss"] < hist[0]["vf_loss"]
    env.close()


def test_ppo2_schedules_are_callables_of_remaining_progress():
    """learning_rate / cliprange as functions of frac = 1 - (update - 1) / n_updates (stable-baselines, ppo2.py:316-331; the
    reference keeps `learning_rate=lr_sch.value` / `linear_schedule(3e-4)` commented in run_moving_docking_ppo2.py:98-99)."""
    n = 1024
    env = qs.make_vec("docking-v2", n, seed=3, randomize="episode")
    seen = []
    algo = ppo.PPO2(env, n_steps=8, nminibatches=2, noptepochs=1, seed=1, fused_policy=True,
                    learning_rate=lambda f: 3e-4 * f, cliprange=lambda f: 0.2 * f)
    algo.learn(4 * n * 8, log_fn=lambda log: seen.append((algo.opt.param_groups[0]["lr"], algo.cliprange, algo.cliprange_vf)))
    fracs = [1.0, 0.75, 0.5, 0.25]
    assert len(seen) == 4
    for (lr, cr, cv), f in zip(seen, fracs):
        assert abs(lr - 3e-4 * f) < 1e-12 and abs(cr - 0.2 * f) < 1e-12 and cv == cr
    env.close()


def test_ppo2_checkpoints_and_load(tmp_path):
    """CheckpointCallback-style saves during learn(), PPO2.load on a fresh env, identical deterministic predictions; the
    archive is the stable-baselines layout (data / parameters / parameter_list)."""
    import zipfile
    n = 2048
    env = qs.make_vec("docking-v2", n, seed=3, randomize="episode")
    algo = ppo.PPO2(env, n_steps=16, nminibatches=4, noptepochs=1, seed=1, fused_policy=True)
    algo.learn(4 * n * 16, save_freq=2 * n * 16, save_path=str(tmp_path))
    saved = sorted((p.name for p in tmp_path.iterdir()), key=lambda s: int(s.split("_")[2]))
    assert saved == ["rl_model_%d_steps.zip" % (2 * n * 16), "rl_model_%d_steps.zip" % (4 * n * 16)], saved
    last = str(tmp_path / saved[-1])
    assert set(zipfile.ZipFile(last).namelist()) == {"data", "parameters", "parameter_list"}
    env2 = qs.make_vec("docking-v2", n, seed=3, randomize="episode")
    again = ppo.PPO2.load(last, env2, fused_policy=True)
    assert again.runner.n_steps == 16 and again.nminibatches == 4
    obs = env2.reset().float()
    assert torch.equal(algo.predict(obs), again.predict(obs))
    for a, b in zip(algo.policy.parameters(), again.policy.parameters()):
        assert torch.equal(a, b)
    env.close()
    env2.close()


def test_ppo2_on_hovering_env_clips_to_its_box():
    """train_drl_hover_ppo2.py's setting: hovering-v0 has 13 observations and actions in [0, 1]^4 (hovering_env.py:36-42); the
    runner must clip to THAT Box before stepping (ppo2.py:478-481) and the torch policy path is used (the tensor-core kernel
    is built for the 12-observation docking policy)."""
    n = 1024
    env = qs.make_vec("hovering-v0", n, seed=1)
    algo = ppo.PPO2(env, n_steps=8, nminibatches=2, noptepochs=1, seed=0)
    assert algo.policy.shared_fc0.in_features == 13 and not algo.runner.in_place
    seen = []
    orig = env.step
    env.step = lambda a: (seen.append(a.clone()), orig(a))[1]
    hist = algo.learn(2 * n * 8)
    assert len(hist) == 2 and all(np.isfinite(h["vf_loss"]) for h in hist)
    lo, hi = min(float(a.min()) for a in seen), max(float(a.max()) for a in seen)
    assert lo >= 0.0 and hi <= 1.0 and lo == 0.0  # a zero-mean Gaussian policy: about half the raw samples are clipped at 0
    with pytest.raises(ValueError):
        ppo.PPO2(env, fused_policy=True)
    env.close()
