"""CPU checks of the kernels' arithmetic: quadsim_b200/csrc/qs_core.cuh compiled for the host (tests/host_harness.cpp,
test-only) replays the reference's golden tapes in float64 (literal path, bar 1e-9 + exact flags) and float32 (fast path,
bar 1e-4 * max(1,|ref|)) and is compared with the fp64 oracle at larger N.  The CUDA build of the same header is what
the `-m gpu` tests exercise through the C ABI."""
import numpy as np
import pytest

from host_harness import HostVecEnv
from oracle import oracle as O
from qs_testutil import golden_params, load_golden


def rel_err(a, ref):
    a = np.asarray(a, dtype=np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(1.0, np.abs(ref)))) if a.size else 0.0


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 1e-4)])
@pytest.mark.parametrize("name", ["dock_v0_kat", "dock_v2_kat", "dock_v0_rand", "dock_v2_rand", "dock_v0_docked", "dock_v0_params", "dock_v2_params"])
def test_reference_tapes(name, dtype, tol):
    """`*_params`: the tapes recorded on the unmodified reference with NON-default mass / inertia / arm length / dt / gravity /
    motor constant / gains / set-points / start positions / dock-port offset, replayed through the RuntimeParams arithmetic
    (what a handle created with use_params = 1 runs on the GPU)."""
    g = load_golden(name)
    env_id = {0: "docking-v0", 1: "docking-v2"}[int(g["kind"])]
    K, Tn = g["actions"].shape[:2]
    seed = int(g["seed"])
    env = HostVecEnv(env_id, K, dtype, seed=max(seed, 0), env_id0=int(g["env_id0"]), randomize=1 if seed >= 0 else 0,
                     half_width=float(g["half_width"]), params=golden_params(g))
    assert rel_err(env.reset(), g["obs0"]) <= (1e-13 if dtype == np.float64 else 1e-6)
    assert np.array_equal(env.get_state()["chaser"][:, :3], g["ini_pos0"])
    for t in range(Tn):
        obs, rew, done, info, term = env.step(g["actions"][:, t])
        assert np.array_equal(done, g["done"][:, t]), t  # these tapes stay clear of thresholds, even in fp32
        assert np.array_equal((info & 4) != 0, g["docking"][:, t] != 0), t
        assert np.array_equal((info & 2) != 0, g["overlimit"][:, t] != 0), t
        assert max(rel_err(obs, g["obs"][:, t]), rel_err(term, g["term_obs"][:, t]), rel_err(rew, g["reward"][:, t])) <= tol, t
        st = env.get_state()
        nd = done == 0
        assert rel_err(st["des_q"], g["des_q"][:, t]) <= tol
        if nd.any():
            assert max(rel_err(st["chaser"][nd], g["chaser"][nd, t]), rel_err(st["target"][nd], g["target"][nd, t]),
                       rel_err(st["u_c"][nd], g["u_c"][nd, t]), rel_err(st["u_t"][nd], g["u_t"][nd, t])) <= tol, t
        for k in np.nonzero(done)[0]:
            assert np.array_equal(st["chaser"][k, :3], g["ini_pos"][k, t])
    env.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("env_id", ["docking-v0", "docking-v2", "hovering-v0"])
def test_runtime_params_with_default_values_equal_the_compiled_in_constants(env_id, dtype):
    """make_runtime_params(qs_default_params) must reproduce DefaultParams: the same trajectories bit for bit (the three
    pre-test thresholds of the attitude clamp are the one deliberate difference, and they only decide whether the exact,
    out-of-line clamp is consulted).  fp64 compares the guarded arithmetic on both sides (runtime parameters never
    speculate)."""
    n, steps = 256, 300
    rand = 2 if env_id.startswith("hover") else 1
    a = HostVecEnv(env_id, n, dtype, seed=4, randomize=rand)
    b = HostVecEnv(env_id, n, dtype, seed=4, randomize=rand, runtime_params=True)
    a.set_speculative(False)
    assert np.array_equal(a.reset(), b.reset())
    rng = np.random.RandomState(0)
    lo = 0.0 if env_id.startswith("hover") else -1.0
    n_done = 0
    for t in range(steps):
        act = rng.uniform(lo, 1, (n, 4)).astype(np.float32)
        ra, rb = a.step(act), b.step(act)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y), t
        n_done += int(ra[2].sum())
    sa, sb = a.get_state(), b.get_state()
    assert all(np.array_equal(sa[k], sb[k]) for k in sa) and (n_done > n or env_id.startswith("hover"))
    a.close(); b.close()


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-9), (np.float32, 1e-4)])
@pytest.mark.parametrize("env_id", ["docking-v0", "docking-v2", "hovering-v0"])
def test_runtime_params_thresholds_vs_oracle(env_id, dtype, tol):
    """The members of qs_env_params_t that are literals inside the reference's step() -- range limit, docking tolerances,
    altitude floor, step limit, attitude clamp angles -- cannot be varied on the live reference; the RuntimeParams arithmetic
    is checked against the parametrised oracle instead (itself pinned to the reference on every attribute-level parameter by
    the *_params tapes): 400 steps, flags exact in fp64."""
    n, steps = 512, 400
    hover = env_id.startswith("hover")
    prm = dict(range_limit=30.0 if hover else (2.5 if env_id == "docking-v0" else 6.0), max_steps=25, min_altitude=0.4, dock_pos_tol=2.1,
               dock_vel_tol=0.6, dock_angle_deg=25.0, att_limit_deg=[70.0, 60.0, 150.0], mass=0.2, dt=0.025)
    rand = 2 if hover else 1
    env = HostVecEnv(env_id, n, dtype, seed=8, randomize=rand, params=prm)
    ora = O.OracleVecEnv(env_id, n, seed=8, randomize=rand, params=prm)
    assert rel_err(env.reset(), ora.reset()) <= (1e-13 if dtype == np.float64 else 1e-6)
    rng = np.random.RandomState(1)
    alive = np.ones(n, bool)
    docked = overtime = 0
    for t in range(steps):
        act = rng.uniform(0.0 if hover else -1.0, 1, (n, 4)).astype(np.float32)
        obs, rew, done, info, term = env.step(act)
        o_obs, o_rew, o_done, o_info, o_term = ora.step(act.astype(np.float64), want_term_obs=True)
        if dtype == np.float64:
            assert np.array_equal(done, o_done) and np.array_equal(info, o_info), t
        alive &= done == o_done
        assert max(rel_err(obs[alive], o_obs[alive]), rel_err(rew[alive], o_rew[alive]), rel_err(term[alive], o_term[alive])) <= tol, t
        docked += int(((o_info & 4) != 0).sum())
        overtime += int(((o_info & 8) != 0).sum())
    assert (~alive).sum() <= 2
    if not hover:
        assert docked > 0 and overtime > 0  # the widened docking tolerances and the 25-step limit were both exercised
    env.close(); ora.close()


def test_f64_scalar_math_polynomials():
    """qs::f64m atan2 / sincos (the polynomial and argument-reduction code shared by host and device; the device-only
    MUFU-seeded reciprocal / rsqrt are covered by tests/test_gpu_parity.py::test_scalar_math_accuracy) vs numpy: <= 3 ulp,
    and the IEEE special cases are the library's."""
    from host_harness import scalar_math
    rng = np.random.RandomState(3)
    n = 1 << 20
    scale = 10.0 ** rng.uniform(-6, 6, n)
    y = rng.uniform(-1, 1, n) * scale
    x = rng.uniform(-1, 1, n) * np.where(rng.rand(n) < 0.5, scale, 1.0)
    y[: n // 4] = rng.uniform(-8, 8, n // 4)
    out = scalar_math(y, x)
    eps = np.finfo(np.float64).eps
    at = np.arctan2(y, x)
    assert (np.abs(out[:, 0] - at) <= 3 * eps * np.abs(at)).all()
    ang = np.abs(y) < 1e4
    for k, ref in ((1, np.sin(y)), (2, np.cos(y))):
        assert (np.abs(out[ang, k] - ref[ang]) <= 3 * eps * np.maximum(np.abs(ref[ang]), 2.0 ** -10)).all()
        assert (np.abs(out[~ang, k] - ref[~ang]) <= 3 * eps).all()
    sy = np.array([0.0, 0.0, -0.0, -0.0, 1.0, -1.0, 0.0, 0.0, np.inf, 1.0, np.inf, -np.inf, 1e-300, 1e300, 5e-324, 1e-320])
    sx = np.array([0.0, -0.0, 0.0, -0.0, 0.0, 0.0, 1.0, -1.0, 1.0, np.inf, np.inf, -np.inf, 1e-300, -1e300, 5e-324, 1.0])
    got, ref = scalar_math(sy, sx)[:, 0], np.arctan2(sy, sx)
    assert np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref))
    assert np.isnan(scalar_math(np.array([np.nan, 1.0]), np.array([1.0, np.nan]))[:, 0]).all()
    ya = np.concatenate([rng.uniform(-1, 1, n // 2), np.sign(rng.uniform(-1, 1, 4096)) * (1 - 10.0 ** rng.uniform(-16, -1, 4096)),
                         [0.0, 0.5, -0.5, 1.0, -1.0, 1e-300, 0.5000000000000001]])
    got, ref = scalar_math(ya, np.ones_like(ya))[:, 5], np.arcsin(ya)
    assert (np.abs(got - ref) <= 3 * eps * np.abs(ref)).all()
    with np.errstate(invalid="ignore"):
        assert np.isnan(scalar_math(np.array([1.0000001, -2.0, np.nan]), np.ones(3))[:, 5]).all()


def test_f64_math_coefficients_are_the_generators():
    """The coefficient tables in qs_core.cuh are exactly what tools/gen_f64_math_coefs.py derives (mpmath Chebyshev fits)."""
    mp = pytest.importorskip("mpmath")
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_f64_math_coefs", os.path.join(root, "tools", "gen_f64_math_coefs.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    src = open(os.path.join(root, "quadsim_b200", "csrc", "qs_core.cuh")).read()
    def table(name):
        body = re.search(r"#define %s \{([^}]*)\}" % name, src).group(1)
        return [float.fromhex(x.strip()) for x in body.split(",")]
    zmax = mp.tan(mp.pi / 8) ** 2 * mp.mpf("1.0001")
    zs = (mp.pi / 4) ** 2 * mp.mpf("1.001")
    assert table("QS_F64M_ATAN") == gen.fit(gen.q_atan, 0, zmax, 11)[0]
    assert table("QS_F64M_SIN") == gen.fit(gen.s_sin, 0, zs, 6)[0]
    assert table("QS_F64M_COS") == gen.fit(gen.c_cos, 0, zs, 6)[0]
    assert table("QS_F64M_ASIN") == gen.fit(gen.r_asin, 0, mp.mpf("0.2501"), 13)[0]


@pytest.mark.parametrize("v", [0, 2])
def test_wild_single_step_fp64(v):
    g = load_golden("dock_v%d_wild" % v)
    n = g["in_t"].shape[0]
    env = HostVecEnv("docking-v%d" % v, n, np.float64)
    env.reset()
    env.set_state(chaser=g["in_chaser"], target=g["in_target"], u_c=g["in_u_c"], u_t=g["in_u_t"], des_q=g["in_des_q"],
                  last_shaping=g["in_last_shaping"], t=g["in_t"])
    obs, rew, done, info, term = env.step(g["in_actions"])
    assert np.array_equal(done, g["out_done"])
    assert np.array_equal((info & 4) != 0, g["out_docking"] != 0) and np.array_equal((info & 2) != 0, g["out_overlimit"] != 0)
    assert rel_err(term, g["out_obs"]) < 1e-10 and rel_err(rew, g["out_reward"]) < 1e-10
    st = env.get_state()
    nd = done == 0
    assert rel_err(st["chaser"][nd], g["out_chaser"][nd]) < 1e-11 and rel_err(st["target"][nd], g["out_target"][nd]) < 1e-11
    assert rel_err(st["u_t"][nd], g["out_u_t"][nd]) < 1e-10 and rel_err(st["des_q"], g["out_des_q"]) < 1e-13


@pytest.mark.parametrize("env_id", ["docking-v0", "docking-v2"])
def test_fp64_speculative_step_equals_guarded_step(env_id):
    """The fp64 kernel's speculative straight-line step (qs::sdouble: range tests replaced by NaN poison, then a replay of
    the guarded step when anything non-finite reached the outputs or the hidden state) must be BIT-identical to the
    guarded step alone -- on flying states (no replays at all) and on wild ones (un-normalised / degenerate quaternions,
    huge rates, zero vectors, gimbal lock, non-finite entries), where the replay has to catch every out-of-range argument."""
    rng = np.random.RandomState(11)
    n = 20000
    def wild_state():
        s = np.zeros((n, 13))
        s[:, :3] = [10, -50, 5] + rng.normal(0, 3, (n, 3))
        s[:, 3:6] = rng.normal(0, 2, (n, 3)) * 10.0 ** rng.uniform(-3, 2, (n, 1))
        q = rng.normal(0, 1, (n, 4)) * 10.0 ** rng.uniform(-1, 0.5, (n, 1))
        s[:, 6:10] = q
        s[:, 10:13] = rng.normal(0, 1, (n, 3)) * 10.0 ** rng.uniform(-2, 3, (n, 1))
        k = n // 20
        s[:k, 6:10] = np.eye(4)[rng.randint(0, 4, k)] * rng.choice([-1, 1, 0.5], (k, 1))  # axis-aligned: exact zeros in the matrices
        s[k:2 * k, 6:10] = 0.5                                                              # w = x = y = z: atan2(0, 0), r12 = 1
        s[2 * k:3 * k, 6:10] = [0.5 ** 0.5, 0.5 ** 0.5, 0, 0]                               # phi = 90 degrees
        s[3 * k:3 * k + 50, rng.randint(0, 13, 50)] = np.nan
        s[3 * k + 50:3 * k + 100, rng.randint(0, 13, 50)] = np.inf
        s[3 * k + 100:4 * k, 6:10] *= 1e-160                                                # squares underflow
        s[4 * k:5 * k, 3:6] = 0.0
        s[4 * k:5 * k, 10:13] = 0.0
        return s
    envs = [HostVecEnv(env_id, n, np.float64, seed=3, randomize=1) for _ in range(2)]
    envs[1].set_speculative(False)
    chaser, target = wild_state(), wild_state()
    des_q = rng.normal(0, 1, (n, 4))
    des_q[: n // 10] = 0.5
    des_q[n // 10: n // 5] = [1, 0, 0, 0]
    u_c, t0 = rng.uniform(0, 1.7, (n, 4)), rng.randint(0, 599, n)
    for e in envs:
        e.reset()
        e.set_state(chaser=chaser, target=target, u_c=u_c, u_t=np.full((n, 4), 0.3), des_q=des_q, t=t0)
    tame = HostVecEnv(env_id, n, np.float64, seed=3, randomize=1)
    tame.reset()
    for t in range(40):
        a = rng.uniform(-1, 1, (n, 4))
        with np.errstate(all="ignore"):
            outs = [e.step(a) for e in envs]
            for x, y in zip(outs[0], outs[1]):
                assert np.array_equal(x, y, equal_nan=True), t
            sa, sb = envs[0].get_state(), envs[1].get_state()
            for k in sa:
                assert np.array_equal(sa[k], sb[k], equal_nan=True), (t, k)
        tame.step(a)
    assert envs[1].replays() == 0 and envs[0].replays() > n // 20  # the wild states do exercise the replay ...
    assert tame.replays() == 0                                       # ... and envs that simply fly (and crash, and reset) never do
    print(env_id, "replayed env-steps:", envs[0].replays(), "of", 40 * n)


@pytest.mark.parametrize("env_id", ["docking-v0", "docking-v2"])
def test_fp64_speculative_step_never_replays_in_flight(env_id):
    """2048 envs x 300 steps of random actions with auto-resets (tumbling, clamping, crashing, re-seating -- everything the
    benchmark workload does): the speculative fp64 step never needs its guarded replay, i.e. the replay is a correctness
    net for exotic states, not a cost of normal operation."""
    rng = np.random.RandomState(2)
    n = 2048
    env = HostVecEnv(env_id, n, np.float64, seed=9, randomize=1)
    env.reset()
    done_total = 0
    for t in range(300):
        done_total += int(env.step(rng.uniform(-1, 1, (n, 4)))[2].sum())
    assert done_total > n and env.replays() == 0


def test_hover_tape():
    g = load_golden("hover_rand")
    K, Tn = g["actions"].shape[:2]
    for dtype, tol in ((np.float64, 1e-9), (np.float32, 1e-4)):
        env = HostVecEnv("hovering-v0", K, dtype, seed=int(g["seed"]), env_id0=int(g["env_id0"]), randomize=2)
        assert rel_err(env.reset(), g["obs0"]) <= (1e-15 if dtype == np.float64 else 1e-6)
        for t in range(Tn):
            obs, rew, done, info, term = env.step(g["actions"][:, t])
            assert np.array_equal(done, g["done"][:, t]), t
            assert max(rel_err(obs, g["obs"][:, t]), rel_err(term, g["state"][:, t]), rel_err(rew, g["reward"][:, t])) <= tol, t


@pytest.mark.parametrize("env_id", ["docking-v0", "docking-v2", "hovering-v0"])
def test_fp32_fast_path_vs_oracle(env_id):
    """256 envs x 600 steps with Philox auto-resets: the fp32 fast path stays within 1e-4*max(1,|ref|) of the fp64 oracle;
    done flags may flip only where the oracle sits within fp32 rounding of a threshold (then that env is dropped)."""
    n, steps, seed = 256, 600, 5
    hover = env_id.startswith("hover")
    rand = 2 if hover else 1
    env = HostVecEnv(env_id, n, np.float32, seed=seed, randomize=rand)
    ora = O.OracleVecEnv(env_id, n, seed=seed, randomize=rand, nthreads=4)
    assert rel_err(env.reset(), ora.reset()) < 1e-6
    rng = np.random.RandomState(0)
    alive = np.ones(n, bool)
    worst = 0.0
    for t in range(steps):
        a = rng.uniform(0 if hover else -1, 1, (n, 4)).astype(np.float32)
        obs, rew, done, info, term = env.step(a)
        o_obs, o_rew, o_done, o_info, o_term = ora.step(a.astype(np.float64), want_term_obs=True)
        alive &= done == o_done
        e = max(rel_err(term[alive], o_term[alive]), rel_err(rew[alive], o_rew[alive]), rel_err(obs[alive], o_obs[alive]))
        worst = max(worst, e)
        assert e <= 1e-4, (t, e)
    assert alive.mean() > 0.98, alive.mean()
    print(env_id, "fp32 worst", worst, "in phase", alive.mean())
