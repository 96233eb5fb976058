"""GPU check of qs_policy_ppo_grad (the PPO2 minibatch gradient on the tensor cores) against torch autograd.

Two references: (1) autograd through a torch emulation of the kernel's precision recipe (bf16 weights and hidden
activations, fp32 accumulation, straight-through rounding) -- the per-sample quantities (losses, d loss / d outputs) must
agree tightly and the parameter gradients to bf16-GEMM accuracy; (2) autograd through the plain fp32 module (what
PPO2.loss does) -- same direction, norm within a few percent.  Then one fused optimisation pass vs the autograd pass."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from quadsim_b200 import ppo
    from quadsim_b200.fused_policy import FusedPolicy


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def _emulated_forward(pol, obs):
    bf = _STE.apply
    h = bf(torch.relu(obs @ bf(pol.shared_fc0.weight).t() + pol.shared_fc0.bias))
    hp = bf(torch.relu(h @ bf(pol.pi_fc0.weight).t() + pol.pi_fc0.bias))
    hv = bf(torch.relu(h @ bf(pol.vf_fc0.weight).t() + pol.vf_fc0.bias))
    return hp @ bf(pol.pi.weight).t() + pol.pi.bias, (hv @ bf(pol.vf.weight).t() + pol.vf.bias).squeeze(-1)


def _loss(algo, forward, mb):
    mean, vpred = forward(mb["obs"])
    neglogp = algo.policy.neglogp(mean, mb["actions"])
    advs = mb["returns"] - mb["values"]
    advs = (advs - advs.mean()) / (advs.std(unbiased=False) + 1e-8)
    vclip = mb["values"] + torch.clamp(vpred - mb["values"], -algo.cliprange_vf, algo.cliprange_vf)
    vf_i = 0.5 * torch.maximum((vpred - mb["returns"]) ** 2, (vclip - mb["returns"]) ** 2)
    ratio = torch.exp(mb["neglogp"] - neglogp)
    pg_i = torch.maximum(-advs * ratio, -advs * torch.clamp(ratio, 1.0 - algo.cliprange, 1.0 + algo.cliprange))
    total = pg_i.mean() - algo.policy.entropy() * algo.ent_coef + vf_i.mean() * algo.vf_coef
    return total, pg_i, vf_i, ratio, neglogp


def _rollout(n=4096, T=12, seed=0):
    env = qs.make_vec("docking-v2", n, seed=seed, randomize="episode", want_info=False)
    algo = ppo.PPO2(env, n_steps=T, nminibatches=4, noptepochs=1, seed=seed, fused_policy=True, ent_coef=0.01)
    with torch.no_grad():
        algo.policy.pi.weight.mul_(20.0)  # means of O(1): clip fractions and value clips are exercised
        algo.policy.logstd.fill_(-0.3)
    algo.fused.sync_weights()
    batch = algo.runner.run()
    with torch.no_grad():  # make the policy differ from the one that acted, as inside an optimisation epoch
        for p in algo.policy.parameters():
            p.add_(0.02 * torch.randn_like(p) * p.abs().mean())
        algo.policy.pi.bias.add_(torch.tensor([0.25, -0.2, 0.1, -0.3], device="cuda"))  # ratios leave [0.8, 1.2]
        algo.policy.vf.bias.add_(0.5)                                                   # value clip (|dv| > 0.2) fires
    algo.fused.sync_weights()
    return env, algo, batch


@pytest.mark.parametrize("B", [1000, 4096 * 3])
def test_ppo_grad_matches_autograd(B):
    env, algo, batch = _rollout()
    M = batch["obs"].shape[0]
    idx = torch.randperm(M, device="cuda")[:B].contiguous()
    grads, logs = algo.fused.ppo_grad(batch, idx, algo.cliprange, algo.cliprange_vf, algo.vf_coef, algo.ent_coef)
    ps = algo.fused._ws["per_sample"][:B]
    mb = {k: batch[k][idx] for k in ("obs", "returns", "actions", "values", "neglogp")}
    params = list(algo.policy.parameters())
    # (1) emulated precision
    mean, vpred = _emulated_forward(algo.policy, mb["obs"])
    mean.retain_grad(); vpred.retain_grad()
    algo.policy.zero_grad()
    total, pg_i, vf_i, ratio, nlp = _loss(algo, lambda o: (mean, vpred), mb)
    total.backward()
    assert float((ps[:, 12] - pg_i).abs().max()) < 2e-2 and float((ps[:, 12] - pg_i).abs().mean()) < 2e-4
    assert float((ps[:, 13] - vf_i).abs().max()) < 2e-2 * max(1.0, float(vf_i.max()))
    assert float((ps[:, :4] - mean.grad).abs().max()) < 0.05 * float(mean.grad.abs().max())
    assert float((ps[:, 4] - vpred.grad).abs().max()) < 0.05 * float(vpred.grad.abs().max())
    assert abs(float(logs["pg_loss"]) - float(pg_i.mean())) < 1e-4 and abs(float(logs["vf_loss"]) - float(vf_i.mean())) < 1e-3 * max(1.0, float(vf_i.mean()))
    assert abs(float(logs["clipfrac"]) - float(((ratio - 1).abs() > algo.cliprange).float().mean())) < 5e-3
    assert abs(float(logs["approxkl"]) - float(0.5 * ((nlp - mb["neglogp"]) ** 2).mean())) < 1e-4
    assert float(logs["clipfrac"]) > 0.01  # the branches are really taken
    for name, prm in algo.policy.named_parameters():
        if prm.grad is None:
            continue  # the unused q head
        a, b = grads[prm].flatten(), prm.grad.flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        assert cos > 0.995 and rel < 0.08, (name, cos, rel)
    # (2) plain fp32 autograd: same direction
    algo.policy.zero_grad()
    _loss(algo, algo.policy.forward, mb)[0].backward()
    for name, prm in algo.policy.named_parameters():
        if prm.grad is None:
            continue
        a, b = grads[prm].flatten(), prm.grad.flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.98 and 0.85 < float(a.norm() / b.norm()) < 1.15, (name, cos)
    env.close()


def test_advantage_statistics_computed_by_the_library():
    """qs_policy_ppo_grad with adv_stats = NULL computes mean and 1 / (std + 1e-8) of (returns - values)[idx] itself
    (ppo2.py:265-266; sums in double, block partials added in a fixed order): same gradients as with statistics handed in
    from a float64 torch reduction, and the per-sample results bit-reproducible from call to call."""
    env, algo, batch = _rollout()
    M = batch["obs"].shape[0]
    idx = torch.randperm(M, device="cuda")[:5000].contiguous()
    args = (algo.cliprange, algo.cliprange_vf, algo.vf_coef, algo.ent_coef)
    snap = lambda gl: ({k: v.clone() for k, v in gl[0].items()}, {k: float(v) for k, v in gl[1].items()})
    g_a, l_a = snap(algo.fused.ppo_grad(batch, idx, *args))
    g_b, l_b = snap(algo.fused.ppo_grad(batch, idx, *args))
    adv = (batch["returns"] - batch["values"])[idx].double()
    stats = torch.stack([adv.mean(), 1.0 / (adv.std(unbiased=False) + 1e-8)]).float()
    g_c, l_c = snap(algo.fused.ppo_grad(batch, idx, *args, adv_stats=stats))
    for prm in g_a:
        scale = float(g_c[prm].abs().max()) + 1e-12
        assert float((g_a[prm] - g_b[prm]).abs().max()) < 1e-5 * scale  # the contraction's split-K adds land in any order
        assert float((g_a[prm] - g_c[prm]).abs().max()) < 2e-4 * scale
    assert abs(l_a["pg_loss"] - l_b["pg_loss"]) < 1e-6 and abs(l_a["pg_loss"] - l_c["pg_loss"]) < 1e-5
    # the statistics themselves: per-sample normalised advantages reproduce from call to call bit for bit
    algo.fused.ppo_grad(batch, idx, *args)
    ps1 = algo.fused._ws["per_sample"][:5000].clone()
    algo.fused.ppo_grad(batch, idx, *args)
    assert torch.equal(ps1, algo.fused._ws["per_sample"][:5000])
    with pytest.raises(ValueError):
        algo.fused.ppo_grad(batch, idx, *args, adv_stats=torch.zeros(3, device="cuda"))
    env.close()


def test_packed_rows_give_the_same_gradient():
    """qs_policy_pack_rollout: one 128-byte row per sample (obs | action | return, value, neglogp); the gradient kernel and the
    advantage statistics gather from it instead of the five arrays -- per-sample results bit-identical."""
    env, algo, batch = _rollout()
    M = batch["obs"].shape[0]
    packed = algo.fused.pack_rollout(batch)
    assert packed.shape == (M, 32)
    assert torch.equal(packed[:, :12], batch["obs"]) and torch.equal(packed[:, 12:16], batch["actions"])
    assert torch.equal(packed[:, 16], batch["returns"]) and torch.equal(packed[:, 17], batch["values"]) and torch.equal(packed[:, 18], batch["neglogp"])
    assert float(packed[:, 19:].abs().max()) == 0.0
    assert algo.fused.pack_rollout(batch, out=packed) is packed  # the buffer is reused
    idx = torch.randperm(M, device="cuda")[:7001].contiguous()
    args = (algo.cliprange, algo.cliprange_vf, algo.vf_coef, algo.ent_coef)
    g_a, l_a = algo.fused.ppo_grad(batch, idx, *args)
    g_a = {k: v.clone() for k, v in g_a.items()}
    ps_a = algo.fused._ws["per_sample"][:7001].clone()
    g_b, l_b = algo.fused.ppo_grad(dict(batch, packed_rows=packed), idx, *args)
    assert torch.equal(ps_a, algo.fused._ws["per_sample"][:7001])
    for prm in g_a:
        scale = float(g_a[prm].abs().max()) + 1e-12
        assert float((g_a[prm] - g_b[prm]).abs().max()) < 1e-5 * scale  # split-K adds land in any order
    with pytest.raises(ValueError):
        algo.fused.ppo_grad(dict(batch, packed_rows=packed[:-1]), idx, *args)
    env.close()


def _unpack_image(img, F, B):
    """An operand image of the weight-gradient contraction (include/quadsim_b200.h: per 128-sample tile, the 16-byte chunk b of
    row r at (b * 16 + (r >> 3)) * 128 + (r & 7) * 16) -> the [B, F] bf16 matrix it holds."""
    tiles = (B + 127) // 128
    x = img[: tiles * F * 256].view(torch.bfloat16).view(tiles, F // 8, 16, 8, 8)  # [tile, b, r >> 3, r & 7, feature in chunk]
    return x.permute(0, 2, 3, 1, 4).reshape(tiles * 128, F)[:B]


@pytest.mark.parametrize("B", [128, 1000, 4096 * 3 + 77])
def test_weight_gradient_contraction_kernel_vs_matmul(B):
    """dw_kernel (tcgen05, both operands MN-major, split-K over CTAs, accumulators in tensor memory, red.global.add) against
    torch.matmul on the very operand rows the gradient kernel left in the workspace: isolates the contraction from the rest
    of the gradient path.  fp32 accumulation of bf16 products on both sides: agreement to accumulation order."""
    env, algo, batch = _rollout()
    idx = torch.randperm(batch["obs"].shape[0], device="cuda")[:B].contiguous()
    algo.fused.ppo_grad(batch, idx, algo.cliprange, algo.cliprange_vf, algo.vf_coef, algo.ent_coef)
    ws = algo.fused._ws
    a1, h, hpv = _unpack_image(ws["a1"], 32, B).float(), _unpack_image(ws["h"], 144, B).float(), _unpack_image(ws["hpv"], 256, B).float()
    dz2, dz1, dout = _unpack_image(ws["dz2"], 256, B).float(), _unpack_image(ws["dz1"], 128, B).float(), _unpack_image(ws["dout"], 16, B).float()
    torch.backends.cuda.matmul.allow_tf32 = False
    for name, got, ref in (("g1", ws["g1"], dz1.t() @ a1), ("g2", ws["g2"], dz2.t() @ h), ("g3", ws["g3"], hpv.t() @ dout)):
        scale = float(ref.abs().max())
        assert scale > 0 and float((got - ref).abs().max()) <= 2e-4 * scale + 1e-7, (name, float((got - ref).abs().max()), scale)
    # the bias block of h really is (1, 1, 0, ...) on every row, the per-sample block's column sums are the kernel's col_sums,
    # and the dOut image is the bf16 rounding of the per-sample output gradients
    assert torch.equal(h[:, 128:130], torch.ones(B, 2, device="cuda")) and float(h[:, 130:].abs().max()) == 0.0
    ps = ws["per_sample"][:B]
    assert torch.allclose(ws["col"], ps.double().sum(0).float(), rtol=2e-4, atol=1e-5)
    assert torch.equal(dout[:, :5], ps[:, :5].to(torch.bfloat16).float()) and float(dout[:, 5:].abs().max()) == 0.0
    env.close()


def test_adam_step_kernel_matches_tensorflow_formula():
    """qs_adam_step: tf.clip_by_global_norm + tf.train.AdamOptimizer(epsilon=1e-5) (ppo2.py:200-207) in one launch, against
    the same formulas in torch float64; the device step counter makes consecutive calls advance the bias correction."""
    torch.manual_seed(0)
    pol = ppo.MlpPolicy().cuda()
    fused = FusedPolicy(pol)
    params = [p for n, p in pol.named_parameters() if not n.startswith("q.")]
    ref_p = [p.detach().double().clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    state, lr, b1, b2, eps, clip = {}, 3e-4, 0.9, 0.999, 1e-5, 0.5
    for step in range(1, 6):
        grads = {p: torch.randn_like(p) * (0.3 if step % 2 else 0.001) for p in params}  # clipped on odd steps, not on even ones
        norm = fused.adam_step(grads, state, lr, clip)
        gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
        assert abs(float(norm) - float(gn)) < 1e-4 * float(gn)
        scale = clip / max(float(gn), clip)
        lr_t = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
        for k, p in enumerate(params):
            g = grads[p].double() * scale
            ref_m[k] += (1 - b1) * (g - ref_m[k])
            ref_v[k] += (1 - b2) * (g * g - ref_v[k])
            ref_p[k] -= lr_t * ref_m[k] / (ref_v[k].sqrt() + eps)
            assert float((p.detach().double() - ref_p[k]).abs().max()) < 2e-6, (step, k)
    assert int(state["step_dev"]) == 5 and state["step"] == 5


@pytest.mark.parametrize("n", [1, 7, 4096, 16 * 1048576, 1000003])
def test_permutation_kernel_is_a_keyed_bijection(n):
    """qs_permutation (the update's minibatch shuffle, ppo2.py:350-352): a bijection of [0, n) for any n, reproducible per seed,
    different across seeds, and well mixed (neighbouring positions land far apart)."""
    pol = ppo.MlpPolicy().cuda()
    fused = FusedPolicy(pol)
    a, b, c = fused.permutation(n, 5), fused.permutation(n, 5), fused.permutation(n, 6)
    assert a.dtype == torch.int64 and torch.equal(a, b)
    assert torch.equal(torch.sort(a).values, torch.arange(n, device="cuda"))
    if n > 100:
        assert not torch.equal(a, c) and float((a == torch.arange(n, device="cuda")).float().mean()) < 0.01
        d = (a[1:] - a[:-1]).abs().double()
        assert float(d.mean()) > 0.2 * n  # a uniform shuffle gives n / 3


def test_fused_update_tracks_autograd_update():
    env, algo, batch = _rollout(n=2048, T=16, seed=2)
    state = {k: v.clone() for k, v in algo.policy.state_dict().items()}
    perms = [torch.randperm(batch["obs"].shape[0], device="cuda") for _ in range(algo.noptepochs)]
    feed = list(perms)
    algo._shuffle = lambda n: feed.pop(0)  # both passes see the same minibatches
    log_f = algo.update(batch)
    after_f = [p.detach().clone() for p in algo.policy.parameters()]
    algo.policy.load_state_dict(state)
    algo.opt = torch.optim.Adam(algo.policy.parameters(), lr=3e-4, eps=1e-5)
    feed = list(perms)
    algo.fused_update = False
    log_a = algo.update(batch)
    moved = [float((a - s).norm()) for a, s in zip(after_f, state.values())]
    for (name, p_a), p_f, s0 in zip(algo.policy.named_parameters(), after_f, state.values()):
        if name.startswith("q."):
            continue
        step_a, step_f = p_a.detach() - s0, p_f - s0
        cos = float(torch.dot(step_a.flatten(), step_f.flatten()) / (step_a.norm() * step_f.norm() + 1e-30))
        # Adam's first steps are sign-like, so direction is what can be compared -- and the two optimisers differ by design: the
        # fused step is TensorFlow's Adam (epsilon outside the bias correction, as the reference trains), the autograd pass torch's
        assert cos > 0.8, (name, cos)
    for k in ("pg_loss", "vf_loss", "approxkl", "clipfrac"):
        assert abs(log_f[k] - log_a[k]) < 0.05 * max(abs(log_a[k]), 0.02), (k, log_f[k], log_a[k])
    assert sum(moved) > 0
    env.close()


def test_graph_captured_update_equals_eager_fused_update():
    """PPO2(graph_update=True): the minibatch step replayed from a CUDA graph must leave the same parameters as the eager
    fused step (same kernels, same order; the capture's warm-up step is undone before the first replay)."""
    results = []
    for graph in (False, True):
        env = qs.make_vec("docking-v2", 2048, seed=5, randomize="episode", want_info=False)
        algo = ppo.PPO2(env, n_steps=16, nminibatches=4, noptepochs=2, seed=3, fused_policy=True, graph_update=graph)
        logs = []
        for _ in range(2):
            batch = algo.runner.run()
            logs.append(algo.update(batch))
        results.append(([p.detach().clone() for p in algo.policy.parameters()], logs))
        env.close()
    (p_eager, l_eager), (p_graph, l_graph) = results
    for a, b in zip(p_eager, p_graph):  # 16 Adam steps of 3e-4 each; the split-K partial sums are added in a different order
        assert torch.allclose(a, b, rtol=1e-4, atol=3e-5), float((a - b).abs().max())
    for le, lg in zip(l_eager, l_graph):
        for k in ("pg_loss", "vf_loss", "approxkl", "clipfrac"):
            assert abs(le[k] - lg[k]) <= 1e-3 * max(1.0, abs(le[k])), (k, le[k], lg[k])
    with pytest.raises(ValueError):
        ppo.PPO2(qs.make_vec("docking-v2", 256), fused_policy=True, graph_update=True, learning_rate=lambda f: 1e-4 * f)
