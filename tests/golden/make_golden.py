#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (sd196821/QuadSim, /root/reference) has no tests and no golden vectors of its own
(SURVEY.md section 4), so parity is pinned on outputs of the reference itself, run here through the import
shim ``oracle/ref_shim.py``.  The GPU box has no /root/reference, hence the vectors are committed.

Every array is float64 (the reference's arithmetic type).  Actions are float32-representable so that the
fp32 kernel, the fp64 kernel, the C oracle and the reference all see bit-identical inputs.

Fixtures
  functions.npz      utils/transform.py + Drone/controller/state2rel leaf functions on random inputs
  sim_pid.npz        run_sim_PID.py:9-54 scenario (BASELINE config 1), 2000 steps
  dock_v{0,2}_kat.npz   constant-action episode (KAT-A), default init, until done
  dock_v{0,2}_rand.npz  K envs x T steps, random actions, SB-style auto-reset, Philox-jittered chaser init
  dock_v0_docked.npz    closed loop with the shipped policy trained_model/best_model_v0.zip (docked + overtime
                        branches), recorded as an action tape
  dock_v{0,2}_wild.npz  single steps from injected wild states (clamp branches, non-unit quaternions, t=599)
  hover_rand.npz     hovering-v0, random actions, auto-reset, Philox-sampled per-env init (once per env)
  policy_v0.npz      weights of the shipped policy best_model_v0.zip + its numpy forward on tape observations
  expert_v0.npz      the PID expert loop of run_expert_record.py:103-160 (generate_PID_expert_traj) on docking-v0, two episodes
                     (crosses a reset, so the scripts' state_des = env.chaser_ini_state aliasing is exercised)
  gae_numpy.npz      the GAE loop of Runner._run (ppo2.py:511-521) transcribed and run by numpy on float32 buffers
  dock_v2_rand64.npz    64 envs x 160 steps of docking-v2 (observation-level record only): the fp32 flag-flip budget against
                        the reference itself needs more than two envs
  dock_v{0,2}_params.npz  tapes with NON-DEFAULT model parameters (mass, inertia, arm length, dt, gravity, motor constant,
                        controller gains, set-points, start positions, dock-port offset) set on the unmodified reference's
                        objects after construction -- pins the EnvParams path of the library (qs_env_params_t)
  clamp_freerun.npz     H2 measured: docking-v2 envs started at attitudes that drive the chaser into the +-85 deg clamp, free
                        running for 50 steps; records when the reference's clamp fired
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

R = ref_shim.load()

# --------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3"), written out
# independently here so that the C oracle and the CUDA kernel are both checked against it.
# Convention (DESIGN.md "RNG"): key = (seed_lo, seed_hi); counter = (env_lo, env_hi, episode, stream);
# u_i = (x_i >> 8) * 2^-24 ; jitter_i = fp32((u_i - 0.5) * scale) ; value = fp32(base + jitter).
# --------------------------------------------------------------------------------------------------
M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = list(ctr)
    k = list(key)
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def philox_jitter3(seed, env_id, episode, stream, scale):
    x = philox4x32_10([env_id & MASK, (env_id >> 32) & MASK, episode & MASK, stream],
                      [seed & MASK, (seed >> 32) & MASK])
    u = np.array([(xi >> 8) for xi in x[:3]], dtype=np.float32) * np.float32(2.0 ** -24)
    return (u - np.float32(0.5)) * np.float32(scale)


def f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


# --------------------------------------------------------------------------------------------------
def gen_functions(rng, n=512):
    T = R.transform
    out = {}
    q = rng.normal(size=(n, 4))
    q[: n // 2] /= np.linalg.norm(q[: n // 2], axis=1, keepdims=True)
    q[: n // 4] *= (1.0 + rng.uniform(-0.05, 0.05, size=(n // 4, 1)))
    out["q"] = q
    out["quat2euler"] = np.array([T.quat2euler(x) for x in q])
    out["quat2rot"] = np.array([T.quat2rot(x) for x in q])
    out["drone_quat2rot"] = np.array([R.Drone.quat2rot(x) for x in q])
    e = rng.uniform(-1, 1, size=(n, 3)) * np.array([np.pi / 2, np.pi, np.pi])
    out["euler"] = e
    out["euler2quat"] = np.array([T.euler2quat(x) for x in e])
    # rot2euler on the products state2rel feeds it (docking_env.py:264-268): R(qb) @ R(qa).T, entries may leave [-1,1]
    q2 = rng.normal(size=(n, 4))
    q2[: n // 2] = q[: n // 2] + 0.05 * rng.normal(size=(n // 2, 4))
    Rab = np.array([T.quat2rot(b) @ T.quat2rot(a).T for a, b in zip(q, q2)])
    out["rot"] = Rab
    with np.errstate(invalid="ignore"):
        out["rot2euler"] = np.array([T.rot2euler(x) for x in Rab])

    d = R.Drone()
    u = np.column_stack([rng.uniform(-1, 9, n), rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), rng.uniform(-1, 1, n)])
    out["u"] = u
    out["u_limit"] = np.array([d.u_limit(x) for x in u])

    s = np.zeros((n, 13))
    s[:, 0:3] = rng.normal(size=(n, 3)) * 3 + np.array([8, -50, 5])
    s[:, 3:6] = rng.normal(size=(n, 3)) * 2
    s[:, 6:10] = q
    s[:, 10:13] = rng.normal(size=(n, 3)) * 3
    out["state"] = s
    out["df"] = np.array([d.df(x, uu) for x, uu in zip(s, out["u_limit"])])
    # Drone.step from injected (state, u_prev) with command u  (quadrotor.py:126-144)
    ucmd = np.column_stack([rng.uniform(-1, 9, n), rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), rng.uniform(-1, 1, n)])
    out["ucmd"] = ucmd
    st_out, u_out = [], []
    for x, up, uc in zip(s, out["u_limit"], ucmd):
        d = R.Drone()
        d.reset(x.copy())
        d.u = up.copy()
        d.integrator = R.RK45(d.f, 0, d.state, d.dt)
        st_out.append(d.step(uc).copy())
        u_out.append(d.u.copy())
    out["step_state"] = np.array(st_out)
    out["step_u"] = np.array(u_out)
    # dock ports (quadrotor.py:213-224) for the two port offsets the envs use
    for name, port in (("c", np.array([0.1, 0.0, 0.0])), ("t", np.array([-0.1, 0.0, 0.0]))):
        pp, pv = [], []
        for x in s:
            d = R.Drone()
            d.reset(x.copy(), port)
            dp = d.get_dock_port_state()
            pp.append(dp["pos"])
            pv.append(dp["vel"])
        out["port_pos_" + name] = np.array(pp)
        out["port_vel_" + name] = np.array(pv)
    # controllers (PIDController.py:179-185, :106-141); des_q is overwritten in place (Q4)
    ctl = R.controller(0.086, 0.18)
    des = np.zeros((n, 13))
    des[:, 0:3] = s[:, 0:3] + rng.normal(size=(n, 3))
    des[:, 3:6] = rng.normal(size=(n, 3)) * 0.3
    des[:, 6:10] = np.array([T.euler2quat(x) for x in rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array([1, 1, 6.0])])
    out["des"] = des
    pid_u, pid_des, vel_u, vel_des = [], [], [], []
    for x, dd in zip(s, des):
        d1 = dd.copy()
        pid_u.append(ctl.PID(d1, x))
        pid_des.append(d1)
        d2 = dd.copy()
        vel_u.append(ctl.vel_controller(d2, x, x))
        vel_des.append(d2)
    out["pid_u"], out["pid_des"] = np.array(pid_u), np.array(pid_des)
    out["vel_u"], out["vel_des"] = np.array(vel_u), np.array(vel_des)
    # state2rel (docking_env.py:257-295)
    s2 = s.copy()
    s2[:, 0:3] += rng.normal(size=(n, 3))
    s2[:, 3:6] = rng.normal(size=(n, 3))
    s2[:, 6:10] = q2
    s2[:, 10:13] = rng.normal(size=(n, 3)) * 2
    out["state_b"] = s2
    rel = []
    for a, b in zip(s, s2):
        da, db = R.Drone(), R.Drone()
        da.reset(a.copy(), np.array([0.1, 0, 0]))
        db.reset(b.copy(), np.array([-0.1, 0, 0]))
        with np.errstate(invalid="ignore"):
            rel.append(R.state2rel(a, b, da.get_dock_port_state(), db.get_dock_port_state()))
    out["state2rel"] = np.array(rel)
    return out


def gen_sim_pid():
    """run_sim_PID.py:9-54 without the plots."""
    T = R.transform
    ini = np.zeros(13)
    ini[6:10] = T.euler2quat(np.array([0.0, 0.0, 0.0]))
    des = np.zeros(13)
    des[0:3] = [-0.2, 0.2, 0.2]
    des[6:10] = T.euler2quat(np.array([0.0, 0.0, 0.0]))
    quad = R.Drone()
    quad.reset(ini)
    ctl = R.controller(quad.get_arm_length(), quad.get_mass())
    n = 2000
    states, us, desq = np.zeros((n, 13)), np.zeros((n, 4)), np.zeros((n, 4))
    for t in range(n):
        now = quad.get_state()
        u = ctl.PID(des, now)
        us[t], states[t], desq[t] = u, now, des[6:10]
        quad.step(u)
    return dict(ini=ini, pos_des=des[0:3].copy(), states=states, u_all=us, des_q=desq, final=quad.get_state().copy())


DOCK_BASE_C = np.array([8.0, -50.0, 5.0])


def apply_params(env, prm):
    """Set non-default model parameters on an UNMODIFIED reference env after construction: everything the reference keeps as
    an attribute (Drone.__init__ quadrotor.py:10-63, controller.__init__ PIDController.py:10-44, DockingEnv.__init__
    docking_env.py:34-99).  Derived attributes are recomputed with the reference's own expressions.  (Thresholds that are
    literals inside step() -- range limit, docking tolerances, 600 steps, the 85/175 degree clamp -- cannot be changed
    without editing the reference; those members of qs_env_params_t are checked against the C oracle only.)"""
    for d in (env.chaser, env.target):
        d.dt, d.gravity, d.mass, d.arm_length = prm["dt"], prm["gravity"], prm["mass"], prm["arm_length"]
        d.Inertia = np.diag(prm["inertia"]).astype(np.float64)
        d.F_max = 4 * d.mass * d.gravity                                    # quadrotor.py:21
        d.motor_lambda = prm["motor_lambda"]                                # quadrotor.py:45
        L, lam = d.arm_length, d.motor_lambda
        d.A = np.array([[0.25, 0, -0.5 / L], [0.25, 0.5 / L, 0], [0.25, 0, 0.5 / L], [0.25, -0.5 / L, 0]])   # :47-50
        d.B = np.array([[1, 1, 1, 1], [0, L, 0, -L], [-L, 0, L, 0]])                                            # :52-54
        d.rotor2control = np.array([[1, 1, 1, 1], [0, L, 0, -L], [-L, 0, L, 0], [lam, -lam, lam, -lam]])       # :56-59
    c = env.target_controller
    c.kp_roll, c.kp_pitch, c.kp_yaw = prm["kp_att"]
    c.kd_roll, c.kd_pitch, c.kd_yaw = prm["kd_att"]
    c.kp_x, c.kp_y, c.kp_z = prm["kp_pos"]
    c.kd_x, c.kd_y, c.kd_z = prm["kd_pos"]
    c.kp_vx, c.kp_vy, c.kp_vz = prm["kp_vel"]
    c.kd_vz = 0.0  # multiplied by (now - last) == 0 in the envs; zeroed so the record cannot depend on it
    c.g, c.mass = prm["gravity"], prm["mass"]                               # PIDController.py:43-44
    env.action_mean = np.ones(4) * env.chaser.mass * env.chaser.gravity / 2.0   # docking_env.py:98-99
    env.action_std = np.ones(4) * env.chaser.mass * env.chaser.gravity / 2.0
    env.chaser_dock_port = np.array([prm["dock_port"], 0.0, 0.0])            # docking_env.py:38
    env.target_dock_port = np.array([-prm["dock_port"], 0.0, 0.0])           # :51
    env.target_ini_state[0:3] = prm["target_ini_pos"]                        # :47
    env.target_state_des[0:3] = prm["target_des_pos"]                        # :59
    env.target_state_des[3:6] = prm["target_des_vel"]                        # moving_docking_env.py:62-66


class DockTape:
    """Drive reference docking envs with an action source; record everything per step."""

    def __init__(self, kind, K, seed=None, env_id0=0, half_width=0.5, params=None, light=False):
        self.kind = kind
        self.cls = R.DockingEnv if kind == 0 else R.MovingDockingEnv
        self.K, self.seed, self.env_id0, self.hw = K, seed, env_id0, half_width
        self.envs = [self.cls() for _ in range(K)]
        self.episode = [0] * K
        self.params, self.light = params, light
        self.base_c = DOCK_BASE_C.copy() if params is None else np.asarray(params["chaser_ini_pos"], dtype=np.float64)
        if params is not None:
            for e in self.envs:
                apply_params(e, params)

    def ini_pos(self, k):
        if self.seed is None:
            return self.base_c.copy()
        j = philox_jitter3(self.seed, self.env_id0 + k, self.episode[k], 0, 2 * self.hw)
        return (self.base_c.astype(np.float32) + j).astype(np.float64)

    def reset(self, k):
        env = self.envs[k]
        p = self.ini_pos(k)
        env.chaser_ini_state[:] = 0.0
        env.chaser_ini_state[0:3] = p
        env.chaser_ini_state[6] = 1.0
        return env.reset().copy(), p

    def run(self, T, policy):
        K = self.K
        rec = dict(actions=np.zeros((K, T, 4)), chaser=np.zeros((K, T, 13)), target=np.zeros((K, T, 13)),
                   u_c=np.zeros((K, T, 4)), u_t=np.zeros((K, T, 4)), des_q=np.zeros((K, T, 4)),
                   term_obs=np.zeros((K, T, 12)), obs=np.zeros((K, T, 12)), reward=np.zeros((K, T)),
                   done=np.zeros((K, T), np.uint8), docking=np.zeros((K, T), np.uint8),
                   overlimit=np.zeros((K, T), np.uint8), ini_pos=np.full((K, T, 3), np.nan),
                   obs0=np.zeros((K, 12)), ini_pos0=np.zeros((K, 3)))
        for k in range(K):
            env = self.envs[k]
            obs, rec["ini_pos0"][k] = self.reset(k)
            rec["obs0"][k] = obs
            for t in range(T):
                a = f32(policy(k, t, obs))
                with np.errstate(invalid="ignore"):
                    o, r, d, info = env.step(a)
                rec["actions"][k, t] = a
                rec["chaser"][k, t], rec["target"][k, t] = env.chaser.state, env.target.state
                rec["u_c"][k, t], rec["u_t"][k, t] = env.chaser.u, env.target.u
                rec["des_q"][k, t] = env.target_state_des[6:10]
                rec["term_obs"][k, t], rec["reward"][k, t], rec["done"][k, t] = o, r, d
                rec["docking"][k, t], rec["overlimit"][k, t] = info["flag_docking"], info["done_overlimit"]
                obs = o.copy()
                if d:  # stable-baselines VecEnv worker: obs <- env.reset()  
                    self.episode[k] += 1
                    obs, rec["ini_pos"][k, t] = self.reset(k)
                rec["obs"][k, t] = obs
        if self.light:
            for k in ("chaser", "target", "u_c", "u_t", "des_q"):
                del rec[k]
            rec["actions"] = rec["actions"].astype(np.float32)  # float32-representable by construction
        rec["kind"] = np.int64(self.kind)
        rec["seed"] = np.int64(-1 if self.seed is None else self.seed)
        rec["env_id0"] = np.int64(self.env_id0)
        rec["half_width"] = np.float64(self.hw)
        return rec


def gen_dock_kat(kind):
    tape = DockTape(kind, 1)
    a = np.array([0.1, -0.2, 0.3, -0.4])
    n = {0: 40, 1: 78}[kind]
    rec = tape.run(n + 5, lambda k, t, o: a)
    assert rec["done"][0, n - 1] == 1 and rec["done"][0, : n - 1].sum() == 0, rec["done"]
    return rec


def gen_dock_rand(kind, K=2, T=1000, seed=1234, env_id0=7):
    rng = np.random.RandomState(100 + kind)
    acts = rng.uniform(-1, 1, size=(K, T, 4)).astype(np.float32)
    tape = DockTape(kind, K, seed=seed, env_id0=env_id0)
    return tape.run(T, lambda k, t, o: acts[k, t])


def gen_dock_rand64(kind=1, K=64, T=160, seed=4321, env_id0=1000):
    rng = np.random.RandomState(300 + kind)
    acts = rng.uniform(-1, 1, size=(K, T, 4)).astype(np.float32)
    tape = DockTape(kind, K, seed=seed, env_id0=env_id0, light=True)
    return tape.run(T, lambda k, t, o: acts[k, t])


# non-default model parameters: every member of qs_env_params_t the reference keeps as an attribute
PARAMS_ALT = dict(dt=0.015, gravity=9.5, mass=0.25, arm_length=0.1, inertia=[0.0003, 0.00028, 0.00045], motor_lambda=0.03,
                  kp_att=[-8.0, -9.0, -7.5], kd_att=[4.0, 4.5, 3.0], kp_pos=[-1.2, -0.9, 40.0], kd_pos=[-1.5, -1.8, 7.0],
                  kp_vel=[-0.6, -0.8, 1.2], chaser_ini_pos=[7.5, -49.0, 6.0], target_ini_pos=[10.5, -49.5, 5.5],
                  target_des_pos=[10.0, -49.0, 5.0], target_des_vel=[0.15, 0.05, 0.0], dock_port=0.12)


def gen_dock_params(kind, K=3, T=250, seed=77, env_id0=11):
    prm = dict(PARAMS_ALT)
    if kind == 0:
        prm["target_des_vel"] = [0.0, 0.0, 0.0]  # docking-v0 keeps des vel 0 (docking_env.py:60-64)
    rng = np.random.RandomState(400 + kind)
    acts = rng.uniform(-1, 1, size=(K, T, 4)).astype(np.float32)
    tape = DockTape(kind, K, seed=seed, env_id0=env_id0, params=prm)
    rec = tape.run(T, lambda k, t, o: acts[k, t])
    for k, v in prm.items():
        rec["param_" + k] = np.asarray(v, dtype=np.float64)
    assert rec["done"].sum() >= K  # every env crosses at least one auto-reset
    return rec


def gen_clamp_freerun(K=160, T=50, seed=5):
    """SURVEY H2, measured: docking-v2 envs whose chaser starts at a large roll / pitch with a body rate pushing it through
    the +-85 degree clamp (quadrotor.py:146-168).  After a clamp the state sits exactly ON the threshold and the reference's
    own next decision is taken by the last ulp of quat2euler(euler2quat(.)); the record keeps, per step, whether the clamp
    fired, so a test can measure how long an independent implementation stays on the reference's trajectory."""
    rng = np.random.RandomState(seed)
    T_ = R.transform
    rec = dict(in_chaser=np.zeros((K, 13)), actions=np.zeros((K, T, 4), np.float32), chaser=np.zeros((K, T, 13)), clamp=np.zeros((K, T), np.uint8),
               done=np.zeros((K, T), np.uint8))
    for k in range(K):
        env = R.MovingDockingEnv()
        sc = np.zeros(13)
        sc[0:3] = [8.0, -50.0, 5.0]
        axis = k % 2
        e = np.zeros(3)
        e[axis] = rng.choice([-1, 1]) * np.deg2rad(rng.uniform(60, 84))
        e[1 - axis] = np.deg2rad(rng.uniform(-20, 20))
        e[2] = np.deg2rad(rng.uniform(-30, 30))
        sc[6:10] = T_.euler2quat(e)
        sc[10 + axis] = -np.sign(e[axis]) * rng.uniform(0.5, 4.0)  # the reference's kinematics (quadrotor.py:98-103): -w grows the angle
        env.chaser_ini_state[:] = sc
        env.reset()
        rec["in_chaser"][k] = sc
        acts = f32(rng.uniform(-0.2, 0.2, size=(T, 4)))
        alive = True
        for t in range(T):
            if alive:
                with np.errstate(invalid="ignore"):
                    y = env.chaser.state + env.chaser.dt * env.chaser.df(env.chaser.state, env.chaser.u)
                    ee = T_.quat2euler(y[6:10])
                    fired = not (abs(ee[0]) <= np.deg2rad(85) and abs(ee[1]) <= np.deg2rad(85) and abs(ee[2]) <= np.deg2rad(175))
                    o, r, d, info = env.step(acts[t])
                rec["clamp"][k, t] = fired
                rec["chaser"][k, t], rec["done"][k, t] = env.chaser.state, d
                if d:
                    alive = False  # no auto-reset here: the record ends with the episode
            rec["actions"][k, t] = acts[t]
    assert rec["clamp"].any(axis=1).mean() > 0.5
    return rec


def gen_dock_docked():
    z = zipfile.ZipFile(os.path.join(ref_shim.REF_ROOT, "trained_model", "best_model_v0.zip"))
    p = np.load(io.BytesIO(z.read("parameters")))
    W = {k: p[k].astype(np.float32) for k in p.files}

    def policy(k, t, obs):
        x = obs.astype(np.float32)
        h = np.maximum(x @ W["model/shared_fc0/w:0"] + W["model/shared_fc0/b:0"], 0)
        h = np.maximum(h @ W["model/pi_fc0/w:0"] + W["model/pi_fc0/b:0"], 0)
        return np.clip(h @ W["model/pi/w:0"] + W["model/pi/b:0"], -1, 1)

    tape = DockTape(0, 1)
    rec = tape.run(640, policy)
    return rec


def gen_dock_wild(kind, n=400, seed=7):
    rng = np.random.RandomState(seed + kind)
    T = R.transform
    cls = R.DockingEnv if kind == 0 else R.MovingDockingEnv
    mg = 0.18 * 9.81
    inp = dict(chaser=np.zeros((n, 13)), target=np.zeros((n, 13)), u_c=np.zeros((n, 4)), u_t=np.zeros((n, 4)),
               des_q=np.zeros((n, 4)), t=np.zeros(n, np.int64), last_shaping=np.zeros(n), actions=np.zeros((n, 4)))
    out = dict(chaser=np.zeros((n, 13)), target=np.zeros((n, 13)), u_c=np.zeros((n, 4)), u_t=np.zeros((n, 4)),
               des_q=np.zeros((n, 4)), obs=np.zeros((n, 12)), reward=np.zeros(n), done=np.zeros(n, np.uint8),
               docking=np.zeros(n, np.uint8), overlimit=np.zeros(n, np.uint8))
    D85, D175 = 85 * np.pi / 180, 175 * np.pi / 180
    i = 0
    while i < n:
        mode = i % 4
        # mode 0: near-nominal (docked regime reachable); 1: moderate; 2: wild attitudes (clamps); 3: far/low (done)
        sc, st = np.zeros(13), np.zeros(13)
        spread = [0.03, 1.0, 1.0, 4.0][mode]
        st[0:3] = np.array([10.0, -50.0, 5.0]) + rng.normal(size=3) * [0.0, 0.3, 0.3, 0.5][mode]
        sc[0:3] = st[0:3] + np.array([-0.2, 0, 0]) + rng.normal(size=3) * spread
        if mode == 3 and rng.rand() < 0.5:
            sc[2] = rng.uniform(0.0, 0.2)
        sc[3:6] = rng.normal(size=3) * [0.02, 0.5, 2.0, 2.0][mode]
        st[3:6] = rng.normal(size=3) * [0.02, 0.3, 1.0, 1.0][mode]
        amp = [0.05, 0.5, 1.0, 0.5][mode]
        ec = rng.uniform(-1, 1, 3) * np.array([1.6, 1.6, 3.2]) * amp
        et = rng.uniform(-1, 1, 3) * np.array([1.6, 1.6, 3.2]) * amp
        sc[6:10] = T.euler2quat(ec) * (1 + rng.uniform(-0.03, 0.03) * (mode > 0))
        st[6:10] = T.euler2quat(et) * (1 + rng.uniform(-0.03, 0.03) * (mode > 0))
        sc[10:13] = rng.normal(size=3) * [0.02, 1.0, 6.0, 1.0][mode]
        st[10:13] = rng.normal(size=3) * [0.02, 1.0, 6.0, 1.0][mode]
        uc = np.array([rng.uniform(0, 4 * mg), rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-1, 1)]) * [0.3, 1, 4, 1][mode]
        ut = np.array([rng.uniform(0, 4 * mg), rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-1, 1)]) * [0.3, 1, 4, 1][mode]
        dq = T.euler2quat(rng.uniform(-1, 1, 3) * np.array([0.4, 0.4, 3.0]) * [0.1, 1, 1, 1][mode])
        t = int(rng.randint(0, 599)) if rng.rand() < 0.9 else 599
        ls = rng.uniform(-6, 0)
        a = f32(rng.uniform(-1, 1, 4))
        # H2: skip states whose post-step euler angles would sit within 1e-9 of a clamp threshold -- checked below
        env = cls()
        env.reset()
        env.chaser.state = sc.copy()
        env.state_chaser = env.chaser.state
        env.target.state = st.copy()
        env.state_target = env.target.state
        env.chaser.u, env.target.u = uc.copy(), ut.copy()
        env.target_state_des[6:10] = dq
        env.t, env.last_shaping = t, ls
        if kind == 1:
            env.info["target"] = env.state_target
        for d in (env.chaser, env.target):
            d.integrator = R.RK45(d.f, 0, d.state, d.dt)
        # threshold-proximity check on the un-clamped Euler prediction
        ok = True
        for d in (env.chaser, env.target):
            y = d.state + d.dt * d.df(d.state, d.u)
            e = T.quat2euler(y[6:10])
            if min(abs(abs(e[0]) - D85), abs(abs(e[1]) - D85), abs(abs(e[2]) - D175)) < 1e-7:
                ok = False
        if not ok:
            continue
        with np.errstate(invalid="ignore", divide="ignore"):
            o, r, dn, info = env.step(a)
        if not (np.all(np.isfinite(o)) and np.isfinite(r)):
            continue
        inp["chaser"][i], inp["target"][i], inp["u_c"][i], inp["u_t"][i] = sc, st, uc, ut
        inp["des_q"][i], inp["t"][i], inp["last_shaping"][i], inp["actions"][i] = dq, t, ls, a
        out["chaser"][i], out["target"][i] = env.chaser.state, env.target.state
        out["u_c"][i], out["u_t"][i], out["des_q"][i] = env.chaser.u, env.target.u, env.target_state_des[6:10]
        out["obs"][i], out["reward"][i], out["done"][i] = o, r, dn
        out["docking"][i], out["overlimit"][i] = info["flag_docking"], info["done_overlimit"]
        i += 1
    rec = {"in_" + k: v for k, v in inp.items()}
    rec.update({"out_" + k: v for k, v in out.items()})
    rec["kind"] = np.int64(kind)
    return rec


def gen_hover_rand(K=2, T=1000, seed=99, env_id0=3):
    T_ = R.transform
    rng = np.random.RandomState(55)
    acts = rng.uniform(0, 1, size=(K, T, 4)).astype(np.float32)
    rec = dict(actions=acts.astype(np.float64), state=np.zeros((K, T, 13)), u=np.zeros((K, T, 4)),
               obs=np.zeros((K, T, 13)), reward=np.zeros((K, T)), done=np.zeros((K, T), np.uint8),
               ini=np.zeros((K, 13)), obs0=np.zeros((K, 13)))
    for k in range(K):
        env = R.HoveringEnv()
        jp = philox_jitter3(seed, env_id0 + k, 0, 0, 2.0)
        je = philox_jitter3(seed, env_id0 + k, 0, 1, 0.4)
        ini = np.zeros(13)
        ini[0:3] = (np.array([0, 0, 5], np.float32) + jp).astype(np.float64)
        ini[6:10] = T_.euler2quat(je.astype(np.float64))
        env.ini_state = ini
        rec["ini"][k] = ini
        obs = env.reset().copy()
        rec["obs0"][k] = obs
        for t in range(T):
            o, r, d, _ = env.step(acts[k, t].astype(np.float64))
            rec["state"][k, t], rec["u"][k, t], rec["reward"][k, t], rec["done"][k, t] = o, env.drone.u, r, d
            obs = o.copy()
            if d:
                obs = env.reset().copy()
            rec["obs"][k, t] = obs
    rec["seed"], rec["env_id0"] = np.int64(seed), np.int64(env_id0)
    return rec


def gen_expert_v0(n_episodes=2):
    """run_expert_record.py:40-45 (set-up) and :103-160 (loop), statement for statement, on docking-v0 (docking-v1 is the same
    env with a once-only np.random position jitter, imitating_docking_env.py:34).

    One deliberate difference, and the literal run kept beside it: the script never clears `info_lst`, so from the second
    episode on `state_last = info_lst[idx_per_epi - 1]['chaser']` (:128) is the FIRST episode's state at that step index and
    vel_controller's kd_vz * (vz_now - vz_last) term (PIDController.py:113, kd_vz = 0.1) is fed a difference between
    episodes.  In the first episode, and in run_expert_policy.py (one episode), state_last is the state the previous step
    returned == the current state, and the term is 0.  The main tape clears info_lst at every reset (term always 0);
    `literal_actions` holds the script's own output: identical through the first step of episode 2, different after."""
    out = _expert_loop(n_episodes, clear_info_lst=True)
    lit = _expert_loop(n_episodes, clear_info_lst=False)
    first = len(out["actions"]) // n_episodes
    assert np.array_equal(out["actions"][:first + 1], lit["actions"][:first + 1]) and not np.array_equal(out["actions"], lit["actions"])
    out["literal_actions"] = lit["actions"]
    return out


def _expert_loop(n_episodes, clear_info_lst):
    env = R.DockingEnv()
    control = R.controller(env.chaser.get_arm_length(), env.chaser.get_mass())
    state_des = env.chaser_ini_state
    kp, kd = 0.35, 0
    rec = dict(actions=[], obs=[], rewards=[], episode_starts=[True], chaser=[], target=[], mem=[], done=[], docking=[])
    episode_returns = np.zeros(n_episodes)
    info_lst = []
    ep_idx, reward_sum, idx_per_epi = 0, 0.0, 0
    obs = env.reset()
    while ep_idx < n_episodes:
        rec["obs"].append(np.array(obs, dtype=np.float64).copy())
        state_last = info_lst[idx_per_epi - 1]["chaser"] if idx_per_epi != 0 else env.chaser_ini_state
        des_vel = kp * (env.state_target[0:3] + np.array([-0.2, 0, 0]) - env.state_chaser[0:3]) + kd * (-env.state_chaser[3:6])
        if idx_per_epi != 0:
            state_des[3:6] = des_vel
        u = control.vel_controller(state_des, env.state_chaser, state_last)
        action = (np.linalg.inv(env.chaser.rotor2control) @ u - env.action_mean) / env.action_std
        obs, reward, done, info = env.step(action)
        rec["actions"].append(action.copy())
        rec["rewards"].append(reward)
        rec["done"].append(done)
        rec["docking"].append(info["flag_docking"])
        rec["chaser"].append(env.chaser.state.copy())
        rec["target"].append(env.target.state.copy())
        rec["mem"].append(state_des[3:10].copy())
        info_lst.append(info)
        rec["episode_starts"].append(done)
        reward_sum += reward
        idx_per_epi += 1
        if done:
            obs = env.reset()
            episode_returns[ep_idx] = reward_sum
            reward_sum, idx_per_epi = 0.0, 0
            ep_idx += 1
            if clear_info_lst:
                info_lst = []
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["episode_starts"] = out["episode_starts"][:-1]
    out["episode_returns"] = episode_returns
    out["done"] = out["done"].astype(np.uint8)
    out["docking"] = out["docking"].astype(np.uint8)
    assert len(out["obs"]) == len(out["actions"]) and out["docking"].sum() > 100
    return out


def gen_gae_numpy():
    """The GAE loop of Runner._run (rl_baselines/ppo2/ppo2.py:503-521) transcribed statement by statement and run by numpy
    on float32 buffers with python-float gamma/lam, exactly as the reference holds them (stable-baselines/TensorFlow are not
    installed, so Runner itself cannot run; this loop has no other dependency)."""
    rng = np.random.RandomState(11)
    n_steps, n = 64, 96
    mb_rewards = rng.normal(size=(n_steps, n)).astype(np.float32)
    mb_values = rng.normal(size=(n_steps, n)).astype(np.float32)
    mb_dones = rng.rand(n_steps, n) < 0.08
    last_values = rng.normal(size=n).astype(np.float32)
    dones = rng.rand(n) < 0.08
    gamma, lam = 0.99, 0.95  # run_moving_docking_ppo2.py:88-104
    mb_advs = np.zeros_like(mb_rewards)
    last_gae_lam = 0
    for step in reversed(range(n_steps)):
        if step == n_steps - 1:
            nextnonterminal = 1.0 - dones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - mb_dones[step + 1]
            nextvalues = mb_values[step + 1]
        delta = mb_rewards[step] + gamma * nextvalues * nextnonterminal - mb_values[step]
        mb_advs[step] = last_gae_lam = delta + gamma * lam * nextnonterminal * last_gae_lam
    mb_returns = mb_advs + mb_values
    return dict(rew=mb_rewards, val=mb_values, dones=mb_dones, last_val=last_values, last_d=dones, gamma=gamma, lam=lam,
                adv=mb_advs, ret=mb_returns)


def gen_policy_v0():
    """Weights of the reference's shipped policy trained_model/best_model_v0.zip (a stable-baselines PPO2 archive) and its
    deterministic actions on the first 64 observations of the docked tape, computed with plain numpy float32."""
    z = zipfile.ZipFile(os.path.join(ref_shim.REF_ROOT, "trained_model", "best_model_v0.zip"))
    p = np.load(io.BytesIO(z.read("parameters")))
    order = json.loads(z.read("parameter_list").decode())
    rec = {"param_order": np.array(order)}
    for k in order:
        rec["param::" + k] = p[k]
    tape = np.load(os.path.join(HERE, "dock_v0_docked.npz"))
    obs = np.concatenate([tape["obs0"], tape["obs"][0, :63]]).astype(np.float32)
    W = {k: p[k].astype(np.float32) for k in p.files}
    h = np.maximum(obs @ W["model/shared_fc0/w:0"] + W["model/shared_fc0/b:0"], 0)
    hp = np.maximum(h @ W["model/pi_fc0/w:0"] + W["model/pi_fc0/b:0"], 0)
    hv = np.maximum(h @ W["model/vf_fc0/w:0"] + W["model/vf_fc0/b:0"], 0)
    rec["obs"], rec["mean"], rec["value"] = obs, hp @ W["model/pi/w:0"] + W["model/pi/b:0"], (hv @ W["model/vf/w:0"] + W["model/vf/b:0"])[:, 0]
    return rec


def main():
    np.random.seed(0)  # HoveringEnv.__init__ draws from the global RNG (hovering_env.py:23-24); overwritten anyway
    jobs = {
        "functions": lambda: gen_functions(np.random.RandomState(1)),
        "sim_pid": gen_sim_pid,
        "dock_v0_kat": lambda: gen_dock_kat(0),
        "dock_v2_kat": lambda: gen_dock_kat(1),
        "dock_v0_rand": lambda: gen_dock_rand(0),
        "dock_v2_rand": lambda: gen_dock_rand(1),
        "dock_v0_docked": gen_dock_docked,
        "dock_v0_wild": lambda: gen_dock_wild(0),
        "dock_v2_wild": lambda: gen_dock_wild(1),
        "hover_rand": gen_hover_rand,
        "expert_v0": gen_expert_v0,
        "gae_numpy": gen_gae_numpy,
        "policy_v0": gen_policy_v0,
        "dock_v2_rand64": gen_dock_rand64,
        "dock_v0_params": lambda: gen_dock_params(0),
        "dock_v2_params": lambda: gen_dock_params(1),
        "clamp_freerun": gen_clamp_freerun,
    }
    only = sys.argv[1:]
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        rec = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print("%-16s %8.1f KB" % (name, os.path.getsize(path) / 1024), flush=True)
    # Philox known answers, produced by the independent implementation above (checked in tests against the
    # Random123 published vectors as well)
    kat = {"philox": [[c, k, philox4x32_10(c, k)] for c, k in (
        ([0, 0, 0, 0], [0, 0]), ([MASK] * 4, [MASK] * 2),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]),
        ([7, 0, 3, 0], [1234, 0]), ([12345678, 1, 99, 1], [0xDEADBEEF, 0x12345678]))]}
    with open(os.path.join(HERE, "philox_kat.json"), "w") as f:
        json.dump(kat, f)


if __name__ == "__main__":
    main()
