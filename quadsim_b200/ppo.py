"""Device-resident PPO2 around the CUDA environments (BASELINE config 5).

The reference trains with stable-baselines' PPO2 (`run_moving_docking_ppo2.py:88-104`; in-tree fork
`rl_baselines/ppo2/ppo2.py`): a numpy `Runner` that round-trips observations/actions through host memory every step
(`ppo2.py:471-498`), GAE in a python loop (`:511-521`) and a TF1 graph for the clipped-surrogate loss (`:138-189`).  Here
observations, actions, rewards, dones, values and advantages never leave the GPU: the env step is the fused CUDA kernel, GAE
is `qs_gae` (one thread per env), the policy MLP and its update use torch (library GEMMs + autograd: this is the *caller* of
the hot path, not the hot path).  Hyper-parameter defaults are the reference's.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _capi
from .distributed import all_reduce_mean_grads
from .sb_zip import load_sb_zip, save_sb_zip


def _ortho_(linear, scale):
    nn.init.orthogonal_(linear.weight, gain=scale)  # SB `linear(..., init_scale)`: orthogonal weights, zero bias
    nn.init.zeros_(linear.bias)


class MlpPolicy(nn.Module):
    """MlpPolicy with net_arch=[128, dict(pi=[128], vf=[128])], act_fun=relu (run_moving_docking_ppo2.py:90-91), diagonal
    Gaussian actions with a state-independent log-std.  Parameter names follow the SB archive (sb_zip.py)."""

    def __init__(self, obs_dim=12, act_dim=4, hidden=128):
        super().__init__()
        self.shared_fc0 = nn.Linear(obs_dim, hidden)
        self.pi_fc0 = nn.Linear(hidden, hidden)
        self.vf_fc0 = nn.Linear(hidden, hidden)
        self.pi = nn.Linear(hidden, act_dim)
        self.vf = nn.Linear(hidden, 1)
        self.q = nn.Linear(hidden, act_dim)  # SB creates it from the vf latent; unused for Gaussian policies, kept for round trips
        self.logstd = nn.Parameter(torch.zeros(1, act_dim))
        for lin in (self.shared_fc0, self.pi_fc0, self.vf_fc0):
            _ortho_(lin, math.sqrt(2))
        _ortho_(self.pi, 0.01)
        _ortho_(self.vf, 1.0)
        _ortho_(self.q, 0.01)

    def forward(self, obs):
        h = torch.relu(self.shared_fc0(obs))
        mean = self.pi(torch.relu(self.pi_fc0(h))).float()  # .float(): no-ops in fp32, widen under bf16 autocast
        value = self.vf(torch.relu(self.vf_fc0(h))).squeeze(-1).float()
        return mean, value

    def neglogp(self, mean, actions):
        # DiagGaussianProbabilityDistribution.neglogp (stable-baselines common/distributions.py)
        std = torch.exp(self.logstd)
        return 0.5 * (((actions - mean) / std) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * actions.shape[-1] + self.logstd.sum()

    def entropy(self):
        return (self.logstd + 0.5 * math.log(2.0 * math.pi * math.e)).sum()

    @torch.no_grad()
    def step(self, obs, deterministic=False, generator=None):
        """model.step: -> actions (unclipped), values, neglogp   (ppo2.py:472)"""
        mean, value = self.forward(obs)
        if deterministic:
            actions = mean
        else:
            noise = torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)
            actions = mean + torch.exp(self.logstd) * noise
        return actions, value, self.neglogp(mean, actions)

    @torch.no_grad()
    def predict(self, obs, deterministic=True, low=-1.0, high=1.0):
        """model.predict as the replay scripts use it (run_trained_moving_docking_ppo2.py:46): clipped to the action Box --
        [-1, 1]^4 for docking (docking_env.py:93); pass low=0 for hovering-v0's [0, 1]^4 (hovering_env.py:42)."""
        return self.step(obs, deterministic=deterministic)[0].clamp(low, high)

    # ---- stable-baselines archive
    _MAP = (("shared_fc0", "model/shared_fc0"), ("pi_fc0", "model/pi_fc0"), ("vf_fc0", "model/vf_fc0"), ("vf", "model/vf"),
            ("pi", "model/pi"), ("q", "model/q"))

    def load_sb_params(self, params):
        with torch.no_grad():
            for attr, name in self._MAP:
                lin = getattr(self, attr)
                if name + "/w:0" in params:
                    lin.weight.copy_(torch.as_tensor(np.asarray(params[name + "/w:0"])).t())
                    lin.bias.copy_(torch.as_tensor(np.asarray(params[name + "/b:0"])))
            if "model/pi/logstd:0" in params:
                self.logstd.copy_(torch.as_tensor(np.asarray(params["model/pi/logstd:0"])))
        return self

    def sb_params(self):
        from collections import OrderedDict
        out = OrderedDict()
        for attr, name in self._MAP[:5]:
            lin = getattr(self, attr)
            out[name + "/w:0"] = lin.weight.detach().t().cpu().numpy().copy()
            out[name + "/b:0"] = lin.bias.detach().cpu().numpy().copy()
        out["model/pi/logstd:0"] = self.logstd.detach().cpu().numpy().copy()
        out["model/q/w:0"] = self.q.weight.detach().t().cpu().numpy().copy()
        out["model/q/b:0"] = self.q.bias.detach().cpu().numpy().copy()
        return out

    @classmethod
    def from_sb_zip(cls, path, device=None):
        _, params = load_sb_zip(path)
        w = params["model/shared_fc0/w:0"]
        pol = cls(obs_dim=w.shape[0], act_dim=params["model/pi/w:0"].shape[1], hidden=w.shape[1]).load_sb_params(params)
        return pol.to(device) if device is not None else pol

    def save_sb_zip(self, path, data=None):
        return save_sb_zip(path, self.sb_params(), data)


def gae(rewards, values, dones, last_values, last_dones, gamma, lam, out=None):
    """qs_gae: Runner._run's advantage loop (ppo2.py:511-521) on [T,N] float32 CUDA tensors -> (advs, returns).
    `out=(advs, returns)` reuses preallocated [T,N] buffers (a training loop should not allocate per rollout)."""
    T, n = rewards.shape
    r, v = rewards.contiguous(), values.contiguous()
    d = dones.to(torch.uint8).contiguous()
    lv, ld = last_values.contiguous(), last_dones.to(torch.uint8).contiguous()
    for t in (r, v, lv):
        if t.dtype != torch.float32 or not t.is_cuda:
            raise ValueError("gae needs float32 CUDA tensors")
    adv, ret = (torch.empty_like(r), torch.empty_like(r)) if out is None else out
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(r.device):
        st = C.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)
        _capi.check(_capi.load().qs_gae(T, n, p(r), p(v), p(d), p(lv), p(ld), float(gamma), float(lam), p(adv), p(ret), st))
    return adv, ret


class DeviceRunner:
    """Runner (ppo2.py:453-527) with every buffer resident on the GPU, time-major [T, N, ...].

    With a FusedPolicy on a float32 docking env the whole step loop is ONE C-ABI call (qs_rollout): the policy kernel and
    the env kernel read and write the rollout rows in place (obs/dones/values have T+1 rows), two launches per step, no
    staging copies.  Otherwise the loop runs here with the torch policy."""

    def __init__(self, env, policy, n_steps, gamma=0.99, lam=0.95, reset_each_rollout=False, generator=None, fused=None, n_slices=2):
        self.env, self.policy, self.n_steps, self.gamma, self.lam = env, policy, int(n_steps), gamma, lam
        self.fused = fused  # FusedPolicy (tcgen05 kernel) or None for the torch forward
        # qs_rollout: 0 = the fused step kernel (one launch per step does model.step and env.step); 1 = two launches per step;
        # 2..8 = env slices on separate streams (one slice's env kernel overlaps another's policy kernel)
        self.n_slices = int(n_slices)
        self.reset_each_rollout = reset_each_rollout  # the in-tree fork's `self.obs = self.env.reset()` (ppo2.py:525)
        self.generator = generator
        n, dev = env.num_envs, env.device
        T = self.n_steps
        f32 = dict(dtype=torch.float32, device=dev)
        self.in_place = fused is not None and env.dtype == torch.float32 and env.obs_dim == 12
        self._obs_all = torch.empty((T + 1, n, env.obs_dim), **f32)
        self._values_all = torch.empty((T + 1, n), **f32)
        self._dones_all = torch.zeros((T + 1, n), dtype=torch.uint8, device=dev)
        self.mb_obs, self.mb_values, self.mb_dones = self._obs_all[:T], self._values_all[:T], self._dones_all[:T]
        self.mb_actions = torch.empty((T, n, 4), **f32)
        self.mb_neglogp = torch.empty((T, n), **f32)
        self.mb_rewards = torch.empty((T, n), **f32)
        self._clipped = torch.empty((n, 4), **f32)
        space = getattr(env, "action_space", None)  # docking: [-1, 1]^4 (docking_env.py:93); hover: [0, 1]^4 (hovering_env.py:42)
        self._act_lo = torch.as_tensor(space.low if space is not None else [-1.0] * 4, **f32)
        self._act_hi = torch.as_tensor(space.high if space is not None else [1.0] * 4, **f32)
        self._advs, self._returns = torch.empty((T, n), **f32), torch.empty((T, n), **f32)
        self.obs = env.reset().float()
        self.dones = torch.zeros(n, dtype=torch.uint8, device=dev)

    def launches_per_rollout(self):
        """Kernel launches of one in-place rollout incl. the bootstrap value pass and GAE."""
        T = self.n_steps
        return T + 2 if self.n_slices == 0 else max(self.n_slices, 1) * (2 * T + 1) + 1

    def _run_in_place(self):
        T, env = self.n_steps, self.env
        self._obs_all[0].copy_(self.obs)      # one 48 B/env copy per ROLLOUT (row T of the previous one), none per step
        self._dones_all[0].copy_(self.dones)
        p = lambda t: t.data_ptr()
        bufs = _capi.RolloutBuffers(p(self._obs_all), p(self.mb_actions), p(self._values_all), p(self.mb_neglogp), p(self.mb_rewards),
                                    p(self._dones_all), p(self._clipped))
        with torch.cuda.device(env.device):
            st = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
            _capi.check(_capi.load().qs_rollout(env._h, self.fused._h, T, C.byref(bufs), self.fused.seed, self.fused._step & 0xFFFFFFFF,
                                                self.n_slices, st))
        self.fused._step += T
        self.obs, self.dones = self._obs_all[T], self._dones_all[T]
        return self._values_all[T]

    @torch.no_grad()
    def run(self):
        env, pol = self.env, self.policy
        if self.in_place:
            last_values = self._run_in_place()
        else:
            for t in range(self.n_steps):
                if self.fused is not None:
                    actions, clipped, values, neglogp = self.fused.step(self.obs)
                    clipped = clipped.to(env.dtype)
                else:
                    actions, values, neglogp = pol.step(self.obs, generator=self.generator)
                    clipped = torch.minimum(torch.maximum(actions, self._act_lo), self._act_hi).to(env.dtype)  # np.clip to the action Box, ppo2.py:478-481
                self.mb_obs[t], self.mb_actions[t], self.mb_values[t], self.mb_neglogp[t] = self.obs, actions, values, neglogp
                self.mb_dones[t] = self.dones
                obs, rew, done, _ = env.step(clipped)
                self.obs = obs.float()
                self.dones = done.to(torch.uint8)
                self.mb_rewards[t] = rew.float()
            last_values = self.fused.step(self.obs, deterministic=True, want_clipped=False)[2] if self.fused is not None else pol.forward(self.obs)[1]
        advs, returns = gae(self.mb_rewards, self.mb_values, self.mb_dones, last_values, self.dones, self.gamma, self.lam,
                            out=(self._advs, self._returns))
        if self.reset_each_rollout:
            self.obs = env.reset().float()
        flat = lambda x: x.reshape((-1,) + tuple(x.shape[2:]))  # time-major flatten; minibatches are shuffled anyway
        return dict(obs=flat(self.mb_obs), returns=flat(returns), actions=flat(self.mb_actions), values=flat(self.mb_values),
                    neglogp=flat(self.mb_neglogp), advs=flat(advs), rewards=self.mb_rewards)


class PPO2:
    """PPO2 with the reference's settings (run_moving_docking_ppo2.py:88-104): n_steps=600, nminibatches=10, noptepochs=10,
    lr=3e-4, gamma=.99, lam=.95, cliprange=.2, ent_coef=0, vf_coef=.5, max_grad_norm=.5, Adam(eps=1e-5).  Loss: ppo2.py:138-189."""

    def __init__(self, env, policy=None, n_steps=600, nminibatches=10, noptepochs=10, learning_rate=3e-4, gamma=0.99, lam=0.95,
                 cliprange=0.2, cliprange_vf=None, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, seed=0, reset_each_rollout=False,
                 fused_policy=False, fused_update=None, graph_update=False, policy_precision="bf16"):
        self.env = env
        self.device = env.device
        torch.manual_seed(seed)
        self.policy = (policy if policy is not None else MlpPolicy(env.obs_dim, 4)).to(self.device)
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        if self.dist:  # every rank starts from rank 0's weights -- BEFORE the fused policy packs its operand images
            for p in self.policy.parameters():
                torch.distributed.broadcast(p.data, 0)
        # the shared `seed` initialises the weights; the action-noise stream of the torch policy is keyed by this rank's
        # first global env id, so ranks do not draw the same noise matrix (the fused kernel keys Philox by env id already)
        self.gen = torch.Generator(device=self.device).manual_seed(int(seed) + int(getattr(env, "env_id0", 0)))
        self.fused = None
        if fused_policy:
            from .fused_policy import FusedPolicy
            # policy_precision="fp32": the rollout forward in plain fp32 (qs_policy_set_precision), e.g. to evaluate a policy
            # exactly as the reference's TensorFlow graph would; the default is the bf16 tensor-core forward
            self.fused = FusedPolicy(self.policy, seed=seed, env_id0=getattr(env, "env_id0", 0), precision=policy_precision)
        # fused_update: the whole minibatch step in hand-written kernels -- qs_policy_ppo_grad (tensor-core forward / backward and the
        # weight-gradient contraction), qs_adam_step (global-norm clip + Adam in TensorFlow's formulation, as the reference
        # trains: ppo2.py:200-207), qs_permutation (the shuffle) -- instead of autograd + torch.optim.Adam + torch.randperm
        self.fused_update = bool(fused_policy) if fused_update is None else bool(fused_update)
        if self.fused_update and self.fused is None:
            raise ValueError("fused_update needs fused_policy=True")
        # graph_update: the whole minibatch step (gradient kernel, weight-gradient GEMMs, clip, Adam, weight re-pack: ~45
        # launches) is captured ONCE in a CUDA graph and replayed; for <= 2^18 envs the update is launch-bound without it.
        # Needs the fused update, constant learning rate / clip ranges and a single process (no collective inside).
        self.graph_update = bool(graph_update)
        if self.graph_update and (not self.fused_update or callable(learning_rate) or callable(cliprange) or callable(cliprange_vf)):
            raise ValueError("graph_update needs fused_update=True and constant learning_rate / cliprange")
        self._graph = None
        self.runner = DeviceRunner(env, self.policy, n_steps, gamma, lam, reset_each_rollout, self.gen, fused=self.fused)
        self.nminibatches, self.noptepochs = nminibatches, noptepochs
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        # learning_rate / cliprange may be callables of the remaining-progress fraction (1 -> 0), as in stable-baselines
        # (ppo2.py:316-331: `frac = 1.0 - (update - 1.0) / n_updates; lr_now = self.learning_rate(frac)`)
        self.learning_rate = learning_rate
        self._cliprange_arg, self._cliprange_vf_arg = cliprange, cliprange_vf
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=learning_rate(1.0) if callable(learning_rate) else learning_rate, eps=1e-5,
                                    )  # the autograd path's optimiser; the fused update has its own (qs_adam_step)
        self._adam = {}  # qs_adam_step state of the fused update (moments, device step counter)
        self._shuffles = 0
        self._seed = int(seed)
        self._set_schedules(1.0)
        self.num_timesteps = 0

    def _set_schedules(self, frac):
        val = lambda x: float(x(frac)) if callable(x) else x
        self.cliprange = val(self._cliprange_arg)
        cv = val(self._cliprange_vf_arg)
        self.cliprange_vf = self.cliprange if cv is None else cv  # legacy default: same as the policy's (ppo2.py:152-156)
        self.lr_now = float(self.learning_rate(frac)) if callable(self.learning_rate) else float(self.learning_rate)
        if hasattr(self, "opt") and callable(self.learning_rate):
            for grp in self.opt.param_groups:
                grp["lr"] = self.lr_now

    def loss(self, mb):
        mean, vpred = self.policy.forward(mb["obs"])
        neglogp = self.policy.neglogp(mean, mb["actions"])
        advs = mb["returns"] - mb["values"]
        advs = (advs - advs.mean()) / (advs.std(unbiased=False) + 1e-8)  # ppo2.py:265-266 (numpy population std)
        old_v, ret = mb["values"], mb["returns"]
        if self.cliprange_vf is not None and self.cliprange_vf >= 0:
            vclip = old_v + torch.clamp(vpred - old_v, -self.cliprange_vf, self.cliprange_vf)
        else:
            vclip = vpred
        vf_loss = 0.5 * torch.maximum((vpred - ret) ** 2, (vclip - ret) ** 2).mean()
        ratio = torch.exp(mb["neglogp"] - neglogp)
        pg_loss = torch.maximum(-advs * ratio, -advs * torch.clamp(ratio, 1.0 - self.cliprange, 1.0 + self.cliprange)).mean()
        entropy = self.policy.entropy()
        total = pg_loss - entropy * self.ent_coef + vf_loss * self.vf_coef
        with torch.no_grad():
            approxkl = 0.5 * ((neglogp - mb["neglogp"]) ** 2).mean()
            clipfrac = ((ratio - 1.0).abs() > self.cliprange).float().mean()
        return total, dict(pg_loss=pg_loss.detach(), vf_loss=vf_loss.detach(), entropy=entropy.detach(), approxkl=approxkl, clipfrac=clipfrac)

    @torch.no_grad()
    def _minibatch_fused(self, batch, idx):
        """One minibatch step in hand-written kernels: gradient (two launches), optimiser (one), weight re-pack (one)."""
        grads, log = self.fused.ppo_grad(batch, idx, self.cliprange, self.cliprange_vf, self.vf_coef, self.ent_coef, want_per_sample=False)
        if self.dist:  # one bucket per minibatch, reduced in place: the four result buffers of the gradient kernels are one allocation
            flat = self.fused._ws["flat"]
            torch.distributed.all_reduce(flat)
            flat /= torch.distributed.get_world_size()
            grads, log = self.fused.grads_from_workspace(self.ent_coef, int(idx.shape[0]))
        self.fused.adam_step(grads, self._adam, self.lr_now, self.max_grad_norm)
        self.fused.sync_weights(sync=False)  # the next minibatch's forward uses the updated weights
        return log

    def _shuffle(self, n_batch):
        """np.random.shuffle(inds) (ppo2.py:350-352).  Fused update: qs_permutation, keyed by (seed, rank, shuffle count);
        otherwise torch.randperm."""
        if self.fused_update:
            self._shuffles += 1
            rank = torch.distributed.get_rank() if self.dist else 0
            return self.fused.permutation(n_batch, (self._seed * 1000003 + rank) * 1000003 + self._shuffles)
        return torch.randperm(n_batch, device=self.device, generator=self.gen)

    def _graph_step(self, batch, idx):
        """Replay the captured minibatch step on `idx` (copied into the graph's static index buffer)."""
        key = (batch["obs"].data_ptr(), batch["returns"].data_ptr(), batch["actions"].data_ptr(), batch["values"].data_ptr(),
               batch["neglogp"].data_ptr(), batch["packed_rows"].data_ptr() if "packed_rows" in batch else 0, idx.shape[0])
        if self._graph is None or self._graph[0] != key:
            if self.dist:
                raise _capi.QuadSimError("graph_update is single-process: the gradient all-reduce is not captured")
            static_idx = idx.clone()
            params = list(self.policy.parameters())
            # One eager step first (outside capture: workspaces and the optimiser's moment tensors are created lazily), then
            # parameters and optimiser state are put back, so the captured step below is the first one that counts.
            snap_p = [prm.detach().clone() for prm in params]
            had_state = "step" in self._adam
            snap_a = {k: {prm: t.clone() for prm, t in self._adam[k].items()} for k in ("m", "v")} if had_state else None
            snap_step = (self._adam["step"], self._adam["step_dev"].clone()) if had_state else None
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side), torch.no_grad():
                self._minibatch_fused(batch, static_idx)
                for prm, p0 in zip(params, snap_p):
                    prm.copy_(p0)
                for k in ("m", "v"):
                    for prm, t in self._adam[k].items():
                        t.copy_(snap_a[k][prm]) if had_state and prm in snap_a[k] else t.zero_()
                if had_state:
                    self._adam["step"] = snap_step[0]
                    self._adam["step_dev"].copy_(snap_step[1])
                else:
                    self._adam["step"] = 0
                    self._adam["step_dev"].zero_()
                self.fused.sync_weights(sync=False)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), torch.no_grad():
                log = self._minibatch_fused(batch, static_idx)
            self._adam["step"] -= 1  # the capture counted a step that has not run yet
            self._graph = (key, graph, static_idx, log)
        _, graph, static_idx, log = self._graph
        static_idx.copy_(idx)
        graph.replay()
        self._adam["step"] += 1  # host mirror of the device step counter the captured optimiser kernel increments
        return {k: v.clone() for k, v in log.items()}

    def update(self, batch):
        n_batch = batch["obs"].shape[0]
        mb_size = n_batch // self.nminibatches
        keys = ("obs", "returns", "actions", "values", "neglogp")
        logs = []
        if self.fused_update or self.graph_update:
            # the rollout as one 128-byte row per sample, once per update: the noptepochs * nminibatches gradient launches gather
            # each sample with one cache line instead of touching five arrays (the buffer's address survives updates: a captured
            # minibatch step reads it)
            self._packed = self.fused.pack_rollout(batch, out=getattr(self, "_packed", None))
            batch = dict(batch, packed_rows=self._packed)
        for _ in range(self.noptepochs):
            perm = self._shuffle(n_batch)
            for start in range(0, mb_size * self.nminibatches, mb_size):
                idx = perm[start:start + mb_size]
                if self.graph_update:
                    logs.append(self._graph_step(batch, idx))
                    continue
                if self.fused_update:
                    logs.append(self._minibatch_fused(batch, idx))
                    continue
                mb = {k: batch[k][idx] for k in keys}
                total, log = self.loss(mb)
                self.opt.zero_grad(set_to_none=True)
                total.backward()
                if self.dist:
                    all_reduce_mean_grads(self.policy.parameters())  # one flat bucket per minibatch
                if self.max_grad_norm is not None:
                    nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)  # tf.clip_by_global_norm
                self.opt.step()
                logs.append(log)
        if self.fused is not None:
            self.fused.sync_weights()  # re-pack the updated parameters into the tensor-core operand images
        return {k: float(torch.stack([l[k] for l in logs]).mean()) for k in logs[0]}

    def learn(self, total_timesteps, log_fn=None, save_freq=None, save_path=None):
        """model.learn(total_timesteps, callback=...) (run_moving_docking_ppo2.py:109).  `save_freq` / `save_path` stand in for
        the reference's CheckpointCallback(save_freq=..., save_path=...) (:79-80): every `save_freq` env steps the policy is
        written as a stable-baselines archive `<save_path>/rl_model_<timesteps>_steps.zip` (rank 0 only)."""
        n_batch = self.env.num_envs * self.runner.n_steps
        n_updates = max(int(total_timesteps) // n_batch, 1)
        history = []
        next_save = self.num_timesteps + save_freq if save_freq else None
        for upd in range(1, n_updates + 1):
            self._set_schedules(1.0 - (upd - 1.0) / n_updates)
            batch = self.runner.run()
            log = self.update(batch)
            self.num_timesteps += n_batch
            log.update(update=upd, timesteps=self.num_timesteps, mean_step_reward=float(batch["rewards"].mean()))
            log.update({"ep_" + k: v for k, v in self.env.stats(all_reduce=self.dist).items() if k in ("mean_return", "mean_length", "episodes", "docked_steps")})
            history.append(log)
            if log_fn:
                log_fn(log)
            if next_save is not None and self.num_timesteps >= next_save and (not self.dist or torch.distributed.get_rank() == 0):
                import os
                os.makedirs(save_path or ".", exist_ok=True)
                self.save(os.path.join(save_path or ".", "rl_model_%d_steps.zip" % self.num_timesteps))
            if next_save is not None:
                while next_save <= self.num_timesteps:
                    next_save += save_freq
        return history

    @classmethod
    def load(cls, path, env, **kw):
        """PPO2.load(path, env=env) (run_trained_moving_docking_ppo2.py:36): the archive's policy weights on a new trainer
        whose hyper-parameters come from the archive's `data` member where it has them (keyword arguments win)."""
        data, _ = load_sb_zip(path)
        for k in ("n_steps", "nminibatches", "noptepochs", "gamma", "lam", "cliprange", "ent_coef", "vf_coef", "max_grad_norm"):
            if k in data and k not in kw and isinstance(data[k], (int, float)):
                kw[k] = data[k]
        algo = cls(env, policy=MlpPolicy.from_sb_zip(path), **kw)
        if algo.fused is not None:
            algo.fused.sync_weights()
        return algo

    def predict(self, obs, deterministic=True):
        """model.predict(obs, deterministic=True)[0] (run_trained_moving_docking_ppo2.py:46): clipped to the action Box."""
        space = getattr(self.env, "action_space", None)
        lo, hi = (float(space.low.min()), float(space.high.max())) if space is not None else (-1.0, 1.0)
        if self.fused is not None and (lo, hi) == (-1.0, 1.0):
            return self.fused.predict(obs, deterministic=deterministic)  # the kernel clips to the docking Box
        return self.policy.predict(obs, deterministic=deterministic, low=lo, high=hi)

    def save(self, path):
        data = dict(gamma=self.runner.gamma, lam=self.runner.lam, n_steps=self.runner.n_steps, nminibatches=self.nminibatches,
                    noptepochs=self.noptepochs, cliprange=self.cliprange, ent_coef=self.ent_coef, vf_coef=self.vf_coef,
                    max_grad_norm=self.max_grad_norm, n_envs=self.env.num_envs)
        return self.policy.save_sb_zip(path, data)
