"""tcgen05 fused forward of the PPO2 MlpPolicy (csrc/qs_policy.cu) behind the C ABI (qs_policy_*).

`FusedPolicy(module)` mirrors `MlpPolicy.step / predict` (quadsim_b200/ppo.py, i.e. model.step / model.predict of the
reference's PPO2, rl_baselines/ppo2/ppo2.py:472) but runs the whole 12 -> 128 -> {128 -> 4, 128 -> 1} network, the Gaussian
sampling and the log-prob in ONE kernel: bf16 tensor-core operands, fp32 accumulation in TMEM, fp32 head and outputs.
The torch module stays the owner of the parameters (autograd / optimiser); call `sync_weights()` after each update."""
import ctypes as C

import torch

from . import _capi


class FusedPolicy:
    def __init__(self, module, seed=0, env_id0=0):
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
        self.sync_weights()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sync_weights(self):
        m = self.module
        ts = [m.shared_fc0.weight, m.shared_fc0.bias, m.pi_fc0.weight, m.pi_fc0.bias, m.vf_fc0.weight, m.vf_fc0.bias,
              m.pi.weight, m.pi.bias, m.vf.weight, m.vf.bias, m.logstd]
        keep = [t.detach().to(torch.float32).contiguous() for t in ts]
        w = _capi.PolicyWeights(*[t.data_ptr() for t in keep])
        with torch.cuda.device(self.device):
            _capi.check(self.lib.qs_policy_set_weights(self._h, C.byref(w), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # `keep` may hold temporaries

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

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.qs_policy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
