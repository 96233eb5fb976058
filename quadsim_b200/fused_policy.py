"""tcgen05 fused forward of the PPO2 MlpPolicy (csrc/qs_policy.cu) behind the C ABI (qs_policy_*).

`FusedPolicy(module)` mirrors `MlpPolicy.step / predict` (quadsim_b200/ppo.py, i.e. model.step / model.predict of the
reference's PPO2, rl_baselines/ppo2/ppo2.py:472) but runs the whole 12 -> 128 -> {128 -> 4, 128 -> 1} network, the Gaussian
sampling and the log-prob in ONE kernel: bf16 tensor-core operands, fp32 accumulation in TMEM, fp32 head and outputs.
The torch module stays the owner of the parameters (autograd / optimiser); call `sync_weights()` after each update."""
import ctypes as C

import torch

from . import _capi


class FusedPolicy:
    def __init__(self, module, seed=0, env_id0=0, precision="bf16"):
        """precision: "bf16" (tensor cores: bf16 weights / hidden activations, fp32 accumulation) or "fp32" (the accurate
        forward on the CUDA cores, ~25x slower: evaluation runs, checks of the fast path).  Either way the same
        (seed, env id, step) draws the same action noise.  The update (ppo_grad) always uses the tensor-core formulation."""
        p = next(module.parameters())
        if not p.is_cuda:
            raise _capi.QuadSimError("FusedPolicy needs the policy on a CUDA device: there is no CPU fallback")
        if module.shared_fc0.in_features != 12 or module.shared_fc0.out_features != 128 or module.pi.out_features != 4:
            raise ValueError("the fused kernel is built for the reference architecture 12 -> 128 -> {128 -> 4, 128 -> 1}")
        self.module, self.device, self.seed, self.env_id0 = module, p.device, int(seed), int(env_id0)
        self.lib = _capi.load()
        h = C.c_void_p()
        _capi.check(self.lib.qs_policy_create(C.byref(h), self.device.index))
        self._h = h
        self._step = 0
        self.precision = None
        self.set_precision(precision)
        self.sync_weights()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_precision(self, precision):
        """"bf16" | "fp32": arithmetic of step / predict and of the two-kernel rollout (qs_policy_set_precision)."""
        modes = {"bf16": _capi.POLICY_BF16, "fp32": _capi.POLICY_FP32}
        if precision not in modes:
            raise ValueError("precision must be 'bf16' or 'fp32', got %r" % (precision,))
        _capi.check(self.lib.qs_policy_set_precision(self._h, modes[precision]))
        self.precision = precision

    def sync_weights(self, sync=True):
        """Re-pack the module's parameters into the kernel's operand images.  sync=False skips the stream synchronise and
        instead keeps the fp32 staging tensors alive until the next call (no host stall inside the update loop)."""
        m = self.module
        ts = [m.shared_fc0.weight, m.shared_fc0.bias, m.pi_fc0.weight, m.pi_fc0.bias, m.vf_fc0.weight, m.vf_fc0.bias,
              m.pi.weight, m.pi.bias, m.vf.weight, m.vf.bias, m.logstd]
        keep = [t.detach().to(torch.float32).contiguous() for t in ts]
        w = _capi.PolicyWeights(*[t.data_ptr() for t in keep])
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_policy_set_weights(self._h, C.byref(w), self._stream()))
            if sync:
                torch.cuda.current_stream(self.device).synchronize()  # `keep` may hold temporaries
            self._keep = keep  # same-stream allocator reuse is ordered behind the pack kernel anyway

    @torch.no_grad()
    def step(self, obs, deterministic=False, step_idx=None, want_clipped=True):
        """-> actions (unclipped), clipped actions (or None), values, neglogp; all float32 [N, ...] on the device."""
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.float().contiguous()
        n = obs.shape[0]
        act = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        clip = torch.empty((n, 4), dtype=torch.float32, device=self.device) if want_clipped else None
        val = torch.empty((n,), dtype=torch.float32, device=self.device)
        nlp = torch.empty((n,), dtype=torch.float32, device=self.device)
        if step_idx is None:
            step_idx = self._step
            self._step += 1
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_policy_step(self._h, n, p(obs), self.seed, self.env_id0, int(step_idx) & 0xFFFFFFFF,
                                                int(bool(deterministic)), p(act), p(clip), p(val), p(nlp), self._stream()))
        return act, clip, val, nlp

    def predict(self, obs, deterministic=True):
        return self.step(obs, deterministic=deterministic)[1]

    # ------------------------------------------------------------------ PPO2 minibatch gradient (qs_policy_ppo_grad)
    def _workspace(self, batch):
        ws = getattr(self, "_ws", None)
        if ws is None or ws["batch"] < batch:
            tiles = (batch + 127) // 128
            raw = lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            f32 = dict(dtype=torch.float32, device=self.device)
            # operand images of the weight-gradient contraction: opaque, sized per 128-sample tile (include/quadsim_b200.h)
            # the four result buffers live in ONE allocation: the multi-GPU update all-reduces it as a single bucket, in place
            flat = torch.empty(128 * 32 + 256 * 144 + 256 * 16 + 16, **f32)
            g1, g2, g3, col = (v.view(s) for v, s in zip(flat.split([4096, 36864, 4096, 16]), ((128, 32), (256, 144), (256, 16), (16,))))
            ws = dict(batch=batch, a1=raw(tiles * 8192), h=raw(tiles * 36864), hpv=raw(tiles * 65536), dz2=raw(tiles * 65536),
                      dz1=raw(tiles * 32768), dout=raw(tiles * 4096), per_sample=torch.empty((batch, 16), **f32),
                      flat=flat, col=col, g1=g1, g2=g2, g3=g3, stats=torch.empty(2, **f32))
            self._ws = ws
        return ws

    @torch.no_grad()
    def ppo_grad(self, batch, idx, cliprange=0.2, cliprange_vf=None, vf_coef=0.5, ent_coef=0.0, want_per_sample=True, adv_stats=None):
        """Gradient of PPO2's loss (ppo2.py:138-189) over the minibatch rows `idx` of the flat rollout `batch` (the dict
        DeviceRunner.run returns), all in hand-written kernels behind ONE C call: forward + per-sample loss + backward
        through the activations on the tensor cores, then the weight-gradient contraction over the whole minibatch
        (split-K, accumulators in tensor memory).  Returns ({param: grad} in the module's layouts, logs).
        adv_stats: optional device tensor (mean, 1 / (std + 1e-8)) of the minibatch's advantages; by default the library
        computes it itself from returns - values over idx (ppo2.py:265-266), in one launch."""
        m, B = self.module, int(idx.shape[0])
        ws = self._workspace(B)
        ret, val = batch["returns"], batch["values"]
        packed = batch.get("packed_rows") if isinstance(batch, dict) else None
        if packed is not None and (packed.dtype != torch.float32 or packed.shape != (batch["obs"].shape[0], 32) or not packed.is_contiguous()):
            raise ValueError("packed_rows must be the [rows, 32] float32 tensor pack_rollout returned for this rollout")
        if adv_stats is not None and (adv_stats.dtype != torch.float32 or adv_stats.numel() != 2 or adv_stats.device != self.device):
            raise ValueError("adv_stats must be two float32 values on %s" % (self.device,))
        p = lambda t: t.data_ptr()
        for k in ("obs", "actions", "returns", "values", "neglogp"):
            t = batch[k]
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError("rollout tensor %r must be contiguous float32 on %s" % (k, self.device))
        if idx.dtype != torch.int64 or not idx.is_contiguous():
            raise ValueError("idx must be a contiguous int64 tensor")
        g = _capi.PPOGrad(p(batch["obs"]), p(batch["actions"]), p(ret), p(val), p(batch["neglogp"]), p(idx), B, p(adv_stats) if adv_stats is not None else None,
                          float(cliprange), float(-1.0 if cliprange_vf is None or cliprange_vf < 0 else cliprange_vf), float(vf_coef),
                          p(ws["a1"]), p(ws["h"]), p(ws["hpv"]), p(ws["dz2"]), p(ws["dz1"]), p(ws["dout"]),
                          p(ws["per_sample"]) if want_per_sample else None, p(ws["col"]), p(ws["g1"]), p(ws["g2"]), p(ws["g3"]),
                          p(packed) if packed is not None else None)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_policy_ppo_grad(self._h, C.byref(g), self._stream()))
        return self.grads_from_workspace(ent_coef, B)

    @torch.no_grad()
    def pack_rollout(self, batch, out=None):
        """The flat rollout as ONE 128-byte row per sample (qs_policy_pack_rollout): `batch["packed_rows"] = pack_rollout(batch)`
        makes every ppo_grad over that rollout gather a sample with one cache line instead of touching five arrays.  Once per
        update; `out` reuses a previous [rows, 32] float32 tensor."""
        rows = int(batch["obs"].shape[0])
        for k in ("obs", "actions", "returns", "values", "neglogp"):
            t = batch[k]
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError("rollout tensor %r must be contiguous float32 on %s" % (k, self.device))
        if out is None or out.shape != (rows, 32) or out.dtype != torch.float32 or out.device != self.device:
            out = torch.empty((rows, 32), dtype=torch.float32, device=self.device)
        p = lambda t: t.data_ptr()
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_policy_pack_rollout(p(batch["obs"]), p(batch["actions"]), p(batch["returns"]), p(batch["values"]),
                                                        p(batch["neglogp"]), rows, p(out), self._stream()))
        return out

    def grads_from_workspace(self, ent_coef, B):
        """({param: grad}, logs) as views of the result buffers the gradient kernels filled (g1, g2, g3, column sums)."""
        m, ws = self.module, self._ws
        col, g1, g2, g3 = ws["col"], ws["g1"], ws["g2"], ws["g3"]
        b2 = g2[:, 128]
        grads = {m.shared_fc0.weight: g1[:, 0:12] + g1[:, 16:28], m.shared_fc0.bias: g1[:, 12],
                 m.pi_fc0.weight: g2[:128, :128], m.pi_fc0.bias: b2[:128], m.vf_fc0.weight: g2[128:, :128], m.vf_fc0.bias: b2[128:],
                 m.pi.weight: g3[:128, 0:4].t(), m.pi.bias: col[0:4], m.vf.weight: g3[128:, 4:5].t(), m.vf.bias: col[4:5],
                 m.logstd: (col[8:12] - ent_coef).view_as(m.logstd)}
        logs = dict(pg_loss=col[12] / B, vf_loss=col[13] / B, approxkl=col[14] / B, clipfrac=col[15] / B, entropy=m.entropy().detach())
        return grads, logs

    # ------------------------------------------------------------------ optimiser step and shuffle (qs_adam_step, qs_permutation)
    def adam_step(self, grads, state, lr, max_grad_norm=0.5, beta1=0.9, beta2=0.999, eps=1e-5):
        """tf.clip_by_global_norm + tf.train.AdamOptimizer(epsilon=1e-5) (ppo2.py:200-207) over `grads` ({param: grad}) in ONE
        launch.  `state` is a dict the caller keeps (moments per parameter, the step count -- held on the device, so that a
        captured CUDA graph of the step replays correctly); the parameters are updated in place."""
        if "step" not in state:
            state.update(step=0, m={}, v={}, norm=torch.zeros(1, dtype=torch.float32, device=self.device),
                         step_dev=torch.zeros(1, dtype=torch.int64, device=self.device))
        params = list(grads.keys())
        gc = [grads[prm] if grads[prm].is_contiguous() else grads[prm].contiguous() for prm in params]
        for prm in params:
            if prm not in state["m"]:
                state["m"][prm] = torch.zeros_like(prm)
                state["v"][prm] = torch.zeros_like(prm)
        tab = (_capi.AdamTensor * len(params))()
        for k, (prm, gk) in enumerate(zip(params, gc)):
            tab[k] = _capi.AdamTensor(prm.data_ptr(), gk.data_ptr(), state["m"][prm].data_ptr(), state["v"][prm].data_ptr(), prm.numel())
        state["step"] += 1
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_adam_step(tab, len(params), float(lr), float(beta1), float(beta2), float(eps), state["step"],
                                              state["step_dev"].data_ptr(), float(max_grad_norm if max_grad_norm is not None else 0.0),
                                              state["norm"].data_ptr(), self._stream()))
        state["keep"] = gc  # the gradient views must outlive the launch
        return state["norm"]

    def permutation(self, n, seed, out=None):
        """np.random.shuffle(inds) of the update (ppo2.py:350-352) as a keyed bijection of [0, n): int64 device tensor."""
        out = torch.empty(n, dtype=torch.int64, device=self.device) if out is None else out
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_permutation(int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, out.data_ptr(), self._stream()))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.qs_policy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
