"""CPU checks of the PPO2 caller side: SB archive reader/writer, the MlpPolicy forward against the reference's
shipped weights evaluated with numpy, the Gaussian log-prob/entropy formulas, and the clipped-surrogate loss against a
numpy transcription of the reference's TF graph (rl_baselines/ppo2/ppo2.py:138-189, :265-266)."""
import math
import os

import numpy as np
import torch

from quadsim_b200.ppo import PPO2, MlpPolicy
from quadsim_b200.sb_zip import load_sb_zip, save_sb_zip
from qs_testutil import load_golden


def _shipped():
    g = load_golden("policy_v0")
    params = {k: g["param::" + k] for k in g["param_order"]}
    return g, params


def test_policy_forward_matches_shipped_weights_numpy():
    g, params = _shipped()
    pol = MlpPolicy().load_sb_params(params)
    mean, val = pol.forward(torch.as_tensor(g["obs"]))
    assert np.max(np.abs(mean.detach().numpy() - g["mean"])) < 5e-6
    assert np.max(np.abs(val.detach().numpy() - g["value"])) < 5e-6
    act = pol.predict(torch.as_tensor(g["obs"]))
    assert np.max(np.abs(act.numpy() - np.clip(g["mean"], -1, 1))) < 5e-6


def test_sb_zip_round_trip(tmp_path):
    g, params = _shipped()
    pol = MlpPolicy().load_sb_params(params)
    path = pol.save_sb_zip(str(tmp_path / "m.zip"), {"gamma": 0.99, "n_envs": 10})
    data, p2 = load_sb_zip(path)
    assert data == {"gamma": 0.99, "n_envs": 10}
    assert list(p2) == list(g["param_order"])  # same variable order as the reference's archive
    for k in params:
        assert np.array_equal(p2[k], params[k]), k
    pol2 = MlpPolicy.from_sb_zip(path)
    x = torch.randn(5, 12)
    assert torch.equal(pol2.forward(x)[0], pol.forward(x)[0])
    save_sb_zip(str(tmp_path / "raw.zip"), {"a/w:0": np.ones((2, 3), np.float32)})
    assert list(load_sb_zip(str(tmp_path / "raw.zip"))[1]) == ["a/w:0"]


def test_gaussian_neglogp_and_entropy():
    torch.manual_seed(0)
    pol = MlpPolicy()
    with torch.no_grad():
        pol.logstd.copy_(torch.tensor([[0.1, -0.2, 0.3, 0.0]]))
    mean, a = torch.randn(7, 4), torch.randn(7, 4)
    ref = -torch.distributions.Normal(mean, pol.logstd.exp().expand_as(mean)).log_prob(a).sum(-1)
    assert torch.allclose(pol.neglogp(mean, a), ref, atol=1e-6)
    ent = torch.distributions.Normal(mean, pol.logstd.exp().expand_as(mean)).entropy().sum(-1)
    assert torch.allclose(pol.entropy().expand(7), ent, atol=1e-6)
    acts, vals, nlp = pol.step(torch.randn(7, 12), deterministic=True)
    assert torch.allclose(nlp, torch.full((7,), 0.5 * math.log(2 * math.pi) * 4 + float(pol.logstd.sum())), atol=1e-6)


class _FakeEnv:
    num_envs, obs_dim, dtype = 4, 12, torch.float32
    device = torch.device("cpu")

    def reset(self):
        return torch.zeros(4, 12)


def test_loss_matches_numpy_transcription_of_tf_graph():
    torch.manual_seed(1)
    algo = PPO2(_FakeEnv(), n_steps=2, cliprange=0.2, ent_coef=0.01, vf_coef=0.5)
    n = 64
    mb = dict(obs=torch.randn(n, 12), actions=torch.randn(n, 4) * 0.5, returns=torch.randn(n), values=torch.randn(n),
              neglogp=torch.randn(n) * 0.1 + 3.7)
    total, log = algo.loss(mb)
    # numpy restatement (float64) of ppo2.py:146-189 with advs normalised as in _train_step (:265-266)
    pol = algo.policy
    mean, vpred = [t.detach().double().numpy() for t in pol.forward(mb["obs"])]
    logstd = pol.logstd.detach().double().numpy()
    a, ret, oldv, oldnlp = [mb[k].double().numpy() for k in ("actions", "returns", "values", "neglogp")]
    neglogpac = 0.5 * np.sum(((a - mean) / np.exp(logstd)) ** 2, axis=-1) + 0.5 * np.log(2 * np.pi) * 4 + logstd.sum()
    entropy = np.sum(logstd + 0.5 * np.log(2 * np.pi * np.e))
    advs = ret - oldv
    advs = (advs - advs.mean()) / (advs.std() + 1e-8)
    vclip = oldv + np.clip(vpred - oldv, -0.2, 0.2)
    vf_loss = 0.5 * np.mean(np.maximum((vpred - ret) ** 2, (vclip - ret) ** 2))
    ratio = np.exp(oldnlp - neglogpac)
    pg_loss = np.mean(np.maximum(-advs * ratio, -advs * np.clip(ratio, 0.8, 1.2)))
    loss = pg_loss - entropy * 0.01 + vf_loss * 0.5
    assert abs(float(total) - loss) < 1e-4 * max(1, abs(loss))
    assert abs(float(log["pg_loss"]) - pg_loss) < 1e-4 and abs(float(log["vf_loss"]) - vf_loss) < 1e-4
    assert abs(float(log["approxkl"]) - 0.5 * np.mean((neglogpac - oldnlp) ** 2)) < 1e-3
    assert abs(float(log["clipfrac"]) - np.mean(np.abs(ratio - 1.0) > 0.2)) < 1e-6
