"""Pin the fp64 CPU oracle (oracle/quadsim_oracle.c) against outputs of the UNMODIFIED reference.

The fixtures in tests/golden/ were produced by tests/golden/make_golden.py, which imports the reference
(sd196821/QuadSim) itself.  The reference has no tests / golden vectors of its own.
Tolerances: leaf functions <= 5e-13 abs (libm vs numpy SVML last-ulp differences, amplified by gains <= 50 and
1/I ~ 4e3); free-running tapes <= 1e-9 abs (BASELINE.json north_star fp64 bar); all flags bit-exact.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from qs_testutil import GOLDEN, close, golden_params, load_golden


@pytest.fixture(scope="module")
def F():
    return load_golden("functions")


def test_philox_published_vectors():
    # Random123 kat_vectors (philox4x32 10 rounds)
    assert list(O.philox4x32_10([0] * 4, [0] * 2)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    kat = json.load(open(os.path.join(GOLDEN, "philox_kat.json")))["philox"]
    for c, k, out in kat:
        assert list(O.philox4x32_10(c, k)) == out


def test_transforms(F):
    close(O.quat2euler(F["q"]), F["quat2euler"], 1e-14)
    close(O.euler2quat(F["euler"]), F["euler2quat"], 1e-15)
    close(O.quat2rot(F["q"]), F["quat2rot"], 1e-14)
    close(O.quat2rot(F["q"]), F["drone_quat2rot"], 1e-14)
    ref = F["rot2euler"]
    got = O.rot2euler(F["rot"])
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    close(np.nan_to_num(got), np.nan_to_num(ref), 1e-14)


def test_plant(F):
    close(O.u_limit(F["u"]), F["u_limit"], 1e-14)
    close(O.df(F["state"], F["u_limit"]), F["df"], 2e-9)  # |ds| up to 1e5 (1/I * wild rates): relative 1e-14
    s, u = O.drone_step(F["state"], F["u_limit"], F["ucmd"])
    close(s, F["step_state"], 5e-11)
    close(u, F["step_u"], 1e-14)
    for name, d in (("c", 0.1), ("t", -0.1)):
        p, v = O.dock_port(F["state"], d)
        close(p, F["port_pos_" + name], 1e-14)
        close(v, F["port_vel_" + name], 1e-14)


def test_controllers(F):
    u, des = O.pid(F["des"], F["state"])
    close(u, F["pid_u"], 5e-13)
    close(des, F["pid_des"], 1e-14)
    u, des = O.vel_controller(F["des"], F["state"], F["state"])
    close(u, F["vel_u"], 5e-13)
    close(des, F["vel_des"], 1e-14)


def test_state2rel(F):
    ref = F["state2rel"]
    got = O.state2rel(F["state"], F["state_b"])
    ok = np.isfinite(ref).all(axis=1)
    assert ok.sum() > 400
    scale = np.maximum(1.0, np.abs(ref[ok]))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-11  # tan/1/cos amplify near phi = +-pi/2


def test_sim_pid_config1():
    """BASELINE config 1: run_sim_PID.py:9-54 (single drone + PID, 2000 steps)."""
    g = load_golden("sim_pid")
    s = g["ini"][None].copy()
    up = np.zeros((1, 4))
    des = np.zeros((1, 13))
    des[0, 0:3] = g["pos_des"]
    des[0, 6] = 1.0
    for t in range(2000):
        assert np.max(np.abs(s[0] - g["states"][t])) < 1e-11, t
        u, des = O.pid(des, s)
        assert np.max(np.abs(u[0] - g["u_all"][t])) < 1e-9, t
        s, up = O.drone_step(s, up, u)
    close(s[0], g["final"], 1e-11)
    close(s[0, :3], [-0.2, 0.2, 0.2], 1e-7)


def _replay_dock(name, tol=1e-9):
    g = load_golden(name)
    kind = {0: O.DOCK_V0, 1: O.DOCK_V2}[int(g["kind"])]
    K, T = g["actions"].shape[:2]
    seed = int(g["seed"])
    light = "chaser" not in g  # observation-level record only (dock_v2_rand64)
    env = O.OracleVecEnv(kind, K, seed=max(seed, 0), env_id0=int(g["env_id0"]),
                         randomize=O.RAND_EPISODE if seed >= 0 else O.RAND_NONE, half_width=float(g["half_width"]),
                         params=golden_params(g))
    obs0 = env.reset()
    close(obs0, g["obs0"], 1e-13)
    close(env.field("chaser")[:, :3], g["ini_pos0"], 0.0)
    worst = 0.0
    for t in range(T):
        obs, rew, done, info, term = env.step(g["actions"][:, t], want_term_obs=True)
        assert np.array_equal(done, g["done"][:, t]), t
        assert np.array_equal((info & O.INFO_DOCKING) != 0, g["docking"][:, t] != 0), t
        assert np.array_equal((info & O.INFO_OVERLIMIT) != 0, g["overlimit"][:, t] != 0), t
        e = max(np.max(np.abs(obs - g["obs"][:, t])), np.max(np.abs(term - g["term_obs"][:, t])),
                np.max(np.abs(rew - g["reward"][:, t])))
        if not light:
            e = max(e, np.max(np.abs(env.field("des_q") - g["des_q"][:, t])))
        nd = done == 0  # after a done the oracle holds the re-seated state, the tape the terminal one
        if nd.any() and not light:
            e = max(e, np.max(np.abs(env.field("chaser")[nd] - g["chaser"][nd, t])),
                    np.max(np.abs(env.field("target")[nd] - g["target"][nd, t])),
                    np.max(np.abs(env.field("u_c")[nd] - g["u_c"][nd, t])),
                    np.max(np.abs(env.field("u_t")[nd] - g["u_t"][nd, t])))
        for k in np.nonzero(done)[0]:
            assert np.array_equal(env.field("chaser")[k, :3], g["ini_pos"][k, t]), (k, t)
        worst = max(worst, e)
        assert e <= tol, (t, e)
    return g, worst


@pytest.mark.parametrize("name", ["dock_v0_kat", "dock_v2_kat", "dock_v0_rand", "dock_v2_rand", "dock_v0_docked", "dock_v2_rand64",
                                  "dock_v0_params", "dock_v2_params"])
def test_dock_tapes(name):
    """`*_params`: tapes recorded with NON-default mass / inertia / arm length / dt / gravity / motor constant / gains /
    set-points / start positions / dock-port offset on the unmodified reference (make_golden.apply_params)."""
    g, worst = _replay_dock(name)
    print(name, "worst abs err", worst)
    if name.endswith("_params"):
        assert golden_params(g)["mass"] == 0.25 and g["done"].sum() >= g["actions"].shape[0]
    if name == "dock_v0_docked":
        assert g["docking"].sum() == 183 and g["done"][0, 599] == 1  # docked + overtime branches exercised


@pytest.mark.parametrize("v", [0, 2])
def test_dock_wild_single_step(v):
    g = load_golden("dock_v%d_wild" % v)
    n = g["in_t"].shape[0]
    env = O.OracleVecEnv(O.DOCK_V0 if v == 0 else O.DOCK_V2, n)
    env.reset()
    for f in ("chaser", "target", "u_c", "u_t", "des_q", "last_shaping", "t"):
        env.field(f)[...] = g["in_" + f]
    obs, rew, done, info, term = env.step(g["in_actions"], want_term_obs=True)
    assert np.array_equal(done, g["out_done"])
    assert np.array_equal((info & O.INFO_DOCKING) != 0, g["out_docking"] != 0)
    assert np.array_equal((info & O.INFO_OVERLIMIT) != 0, g["out_overlimit"] != 0)
    scale = np.maximum(1.0, np.abs(g["out_obs"]))
    assert np.max(np.abs(term - g["out_obs"]) / scale) < 1e-10
    # obs/reward reach 1e17 where rot2euler takes its |R12| >= 1 branch (phi = +-pi/2, 1/cos(phi) ~ 1.6e16)
    assert np.max(np.abs(rew - g["out_reward"]) / np.maximum(1.0, np.abs(g["out_reward"]))) < 1e-10
    close(env.field("des_q"), g["out_des_q"], 1e-13)
    nd = done == 0
    close(env.field("chaser")[nd], g["out_chaser"][nd], 1e-11)
    close(env.field("target")[nd], g["out_target"][nd], 1e-11)
    close(env.field("u_c")[nd], g["out_u_c"][nd], 1e-12)
    close(env.field("u_t")[nd], g["out_u_t"][nd], 1e-10)
    assert done.sum() > 20 and (info & O.INFO_DOCKING).astype(bool).sum() > 10


def test_clamp_freerun_h2_measured():
    """SURVEY H2, measured instead of argued.  160 docking-v2 envs are driven into the +-85 degree attitude clamp
    (quadrotor.py:146-168).  Until an env's first clamp the oracle must sit on the reference's trajectory (1e-9); from then
    on the clamped state lies exactly ON the threshold and the reference's own next decision is taken by the last ulp of
    numpy's arctan2/arcsin (SVML) -- the fraction of envs that leave the reference's trajectory afterwards is reported, and
    must be what makes H2 a limitation: neither zero (then there would be nothing to document) nor everything."""
    g = load_golden("clamp_freerun")
    K, T = g["actions"].shape[:2]
    env = O.OracleVecEnv(O.DOCK_V2, K, auto_reset=False)
    env.reset()
    env.field("chaser")[...] = g["in_chaser"]
    first = np.where(g["clamp"].any(axis=1), g["clamp"].argmax(axis=1), T)
    alive = np.ones(K, bool)
    diverged_at = np.full(K, -1)
    for t in range(T):
        env.step(g["actions"][:, t].astype(np.float64))
        err = np.max(np.abs(env.field("chaser") - g["chaser"][:, t]), axis=1)
        pre = alive & (t < first)
        assert err[pre].max(initial=0.0) <= 1e-9, (t, err[pre].max())
        new = alive & (t >= first) & (err > 1e-6) & (diverged_at < 0)
        diverged_at[new] = t
        alive &= g["done"][:, t] == 0
    clamped = first < T
    frac = float((diverged_at >= 0)[clamped].mean())
    print("H2: %d of %d envs clamp; %.1f %% of those leave the reference's trajectory within %d steps (first clamp at step %.1f on average)"
          % (clamped.sum(), K, 100 * frac, T, first[clamped].mean()))
    assert clamped.sum() >= K // 2
    assert 0.0 <= frac < 0.9


def test_hover_tape():
    g = load_golden("hover_rand")
    K, T = g["actions"].shape[:2]
    env = O.OracleVecEnv(O.HOVER, K, seed=int(g["seed"]), env_id0=int(g["env_id0"]), randomize=O.RAND_ONCE)
    obs0 = env.reset()
    close(obs0, g["obs0"], 1e-15)
    for t in range(T):
        obs, rew, done, info, term = env.step(g["actions"][:, t], want_term_obs=True)
        assert np.array_equal(done, g["done"][:, t]), t
        close(obs, g["obs"][:, t], 1e-9)
        close(term, g["state"][:, t], 1e-9)
        close(rew, g["reward"][:, t], 1e-9)
    assert g["done"].sum() >= 4


def test_gae_matches_reference_loop_bit_exact():
    """Runner._run's GAE loop (rl_baselines/ppo2/ppo2.py:511-521), including numpy's float32/float64 promotion."""
    g = load_golden("gae_numpy")
    adv, ret = O.gae(g["rew"], g["val"], g["dones"], g["last_val"], g["last_d"], float(g["gamma"]), float(g["lam"]))
    assert np.array_equal(adv, g["adv"]) and np.array_equal(ret, g["ret"])
    assert adv.dtype == np.float32 and g["ret"].dtype == np.float32
