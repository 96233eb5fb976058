"""qs_step_n (T steps in one launch, state in registers across steps) against T single-step launches and against the fp64
oracle.  The two kernels inline the same arithmetic (qs_core.cuh) but are separate compilations, so nvcc's FMA contraction
may differ in the last bit: fp64 must agree to 1e-9 with identical done flags (the parity bar of the single-step kernel),
fp32 to 1e-4 * max(1, |x|) on every env whose done flags have agreed so far (a last-bit difference can move an env across a
threshold; at most a handful may).  All envs, both dtypes, ragged batch sizes, two consecutive calls; also pinned directly on
the reference tape of BASELINE config 2's env (hover_rand)."""
import numpy as np
import pytest
import torch

from qs_testutil import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from oracle import oracle as O


def _rel(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs() / b.abs().clamp(min=1.0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("env_id,n,T", [("docking-v2", 1000, 200), ("docking-v0", 129, 90), ("hovering-v0", 4096, 150), ("hovering-v0", 77, 40)])
def test_step_many_equals_single_steps(env_id, n, T, dtype):
    kw = dict(dtype=dtype, seed=11, randomize="episode" if env_id != "hovering-v0" else None)
    a, b = qs.make_vec(env_id, n, **kw), qs.make_vec(env_id, n, **kw)
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.rand((T, n, 4), device="cuda", generator=g, dtype=torch.float32).to(dtype)
    if env_id != "hovering-v0":
        acts = acts * 2 - 1
    else:
        acts = acts * 1.6  # over-thrust: some hover envs fly out of the 100 m box and auto-reset
    obs_m, rew_m, done_m = a.step_many(acts[:T // 2])
    obs_m2, rew_m2, done_m2 = a.step_many(acts[T // 2:])  # a second call continues from the stored state
    obs_m, rew_m, done_m = torch.cat([obs_m, obs_m2]), torch.cat([rew_m, rew_m2]), torch.cat([done_m, done_m2])
    tol = 1e-9 if dtype == torch.float64 else 1e-4
    alive = torch.ones(n, dtype=torch.bool, device="cuda")
    for t in range(T):
        obs, rew, done, _ = b.step(acts[t])
        alive &= done == done_m[t]
        assert float(_rel(obs_m[t], obs)[alive].max()) <= tol and float(_rel(rew_m[t], rew)[alive].max()) <= tol, t
    flips = int((~alive).sum())
    assert flips == 0 if dtype == torch.float64 else flips <= max(2, n // 500), flips
    sa, sb = a.get_state(), b.get_state()
    assert torch.equal(sa["t"][alive], sb["t"][alive]) and torch.equal(sa["episode"][alive], sb["episode"][alive])
    assert float(_rel(sa["chaser"], sb["chaser"])[alive].max()) <= tol
    st_a, st_b = a.stats(), b.stats()
    assert abs(st_a["episodes"] - st_b["episodes"]) <= flips and abs(st_a["length_sum"] - st_b["length_sum"]) <= 600 * flips
    if env_id.startswith("docking"):
        assert int(done_m.sum()) > 0
    a.close()
    b.close()


@pytest.mark.parametrize("env_id", ["docking-v2", "docking-v0", "hovering-v0"])
def test_step_many_fp64_vs_oracle(env_id):
    n, T = 300, 160
    rnd = O.RAND_ONCE if env_id == "hovering-v0" else O.RAND_EPISODE
    env = qs.make_vec(env_id, n, dtype=torch.float64, seed=4, randomize="once" if env_id == "hovering-v0" else "episode")
    ora = O.OracleVecEnv(env_id, n, seed=4, randomize=rnd)
    assert np.max(np.abs(env.reset().cpu().numpy() - ora.reset())) < 1e-12
    rng = np.random.RandomState(3)
    acts = rng.uniform(0 if env_id == "hovering-v0" else -1, 1, size=(T, n, 4)).astype(np.float32).astype(np.float64)
    obs, rew, done = env.step_many(torch.as_tensor(acts, device="cuda"))
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    for t in range(T):
        o_obs, o_rew, o_done, _ = ora.step(acts[t])
        assert np.array_equal(done[t], o_done != 0), t
        assert np.max(np.abs(obs[t] - o_obs) / np.maximum(1, np.abs(o_obs))) < 1e-9 and np.max(np.abs(rew[t] - o_rew) / np.maximum(1, np.abs(o_rew))) < 1e-9, t
    assert np.array_equal(env.stats()["episodes"], ora.stats()[2])
    env.close()
    ora.close()


def test_step_many_replays_reference_hover_tape():
    g = load_golden("hover_rand")
    K, Tn = g["actions"].shape[:2]
    env = qs.make_vec("hovering-v0", K, dtype=torch.float64, seed=int(g["seed"]), env_id0=int(g["env_id0"]))
    env.reset()
    acts = torch.as_tensor(np.ascontiguousarray(g["actions"].transpose(1, 0, 2)), device="cuda")
    obs, rew, done = env.step_many(acts)
    assert np.array_equal(done.cpu().numpy().T, g["done"] != 0)
    rerr = lambda x, ref: float(np.max(np.abs(x - ref) / np.maximum(1.0, np.abs(ref))))
    assert rerr(rew.cpu().numpy().T, g["reward"]) < 1e-9
    assert rerr(obs.cpu().numpy().transpose(1, 0, 2), g["obs"]) < 1e-9
    env.close()
