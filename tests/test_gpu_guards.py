"""Out-of-bounds write guards for the kernels behind the newer C-ABI entry points (compute-sanitizer is not available on the
GPU pool): every output buffer is carved out of a larger allocation filled with a sentinel, the kernels run on ragged sizes
(not multiples of the 128-env tile), and the sentinel bands on both sides must come back untouched."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from quadsim_b200 import _capi, ppo
    from quadsim_b200.fused_policy import FusedPolicy

PAD = 257  # rows of sentinel on each side


class Guarded:
    def __init__(self, rows, cols, dtype, pad=PAD):
        self.sentinel = 123 if dtype == torch.uint8 else (-7777 if dtype in (torch.int32, torch.int64) else -12288.0)  # exactly representable in bfloat16
        self.pad = pad  # (the default leaves byte buffers such as `done` at an odd address on purpose)
        shape = (rows + 2 * pad,) + ((cols,) if cols else ())
        self.full = torch.full(shape, self.sentinel, dtype=dtype, device="cuda")
        self.view = self.full[pad:pad + rows]
        self.rows = rows

    @property
    def ptr(self):
        return C.c_void_p(self.view.data_ptr())

    def intact(self):
        lo, hi = self.full[:self.pad], self.full[self.pad + self.rows:]
        if self.full.dtype == torch.bfloat16:
            lo, hi = lo.float(), hi.float()
        return bool((lo == self.sentinel).all()) and bool((hi == self.sentinel).all())

    def written(self):
        v = self.view.float() if self.view.dtype == torch.bfloat16 else self.view
        return bool((v != self.sentinel).any())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("n", [1, 127, 129, 1000, 148 * 256 + 77])
def test_policy_step_and_ppo_grad_stay_in_bounds(n):
    torch.manual_seed(0)
    pol = ppo.MlpPolicy().cuda()
    fused = FusedPolicy(pol)
    L = fused.lib
    obs = torch.randn(n, 12, device="cuda")
    f32, bf16 = torch.float32, torch.bfloat16
    act, clip, val, nlp = Guarded(n, 4, f32), Guarded(n, 4, f32), Guarded(n, 0, f32), Guarded(n, 0, f32)
    _capi.check(L.qs_policy_step(fused._h, n, C.c_void_p(obs.data_ptr()), 1, 0, 0, 0, act.ptr, clip.ptr, val.ptr, nlp.ptr, _stream()))
    torch.cuda.synchronize()
    for g in (act, clip, val, nlp):
        assert g.intact() and g.written()
    # the fp32 forward (qs_policy_set_precision): 64-row tiles, a ragged last tile, an exactly sized obs allocation before it
    fused.set_precision("fp32")
    obs_g = Guarded(n, 12, f32)
    obs_g.view.copy_(obs)
    act2, clip2, val2, nlp2 = Guarded(n, 4, f32), Guarded(n, 4, f32), Guarded(n, 0, f32), Guarded(n, 0, f32)
    _capi.check(L.qs_policy_step(fused._h, n, obs_g.ptr, 1, 0, 0, 0, act2.ptr, clip2.ptr, val2.ptr, nlp2.ptr, _stream()))
    torch.cuda.synchronize()
    for g in (obs_g, act2, clip2, val2, nlp2):
        assert g.intact() and g.written()
    assert float((act2.view - act.view).abs().max()) < 5e-2 and torch.equal(nlp2.view, nlp.view)  # same draws in both precisions
    fused.set_precision("bf16")
    # minibatch gradient: B rows gathered from a rollout of n rows through idx
    B = max(1, (2 * n) // 3)
    idx = torch.randperm(n, device="cuda")[:B].contiguous()
    ret, oldv, oldnlp = torch.randn(n, device="cuda"), torch.randn(n, device="cuda"), nlp.view.clone()
    stats = torch.tensor([0.0, 1.0], device="cuda")
    tiles = (B + 127) // 128  # the operand images of the weight-gradient contraction are sized per 128-sample tile (header)
    img = lambda F: Guarded(tiles * F * 256, 0, torch.uint8, pad=4096)  # operand images must be 16-byte aligned
    bufs = dict(a1=img(32), h=img(144), hpv=img(256), dz2=img(256), dz1=img(128), dout=img(16), per_sample=Guarded(B, 16, f32),
                col_sums=Guarded(16, 0, f32), g1=Guarded(128, 32, f32), g2=Guarded(256, 144, f32), g3=Guarded(256, 16, f32))
    p = lambda t: t.data_ptr()
    g = _capi.PPOGrad(p(obs), p(act.view), p(ret), p(oldv), p(oldnlp), p(idx), B, p(stats), 0.2, 0.2, 0.5,
                      *[bufs[k].view.data_ptr() for k in ("a1", "h", "hpv", "dz2", "dz1", "dout", "per_sample", "col_sums", "g1", "g2", "g3")])
    _capi.check(L.qs_policy_ppo_grad(fused._h, C.byref(g), _stream()))
    torch.cuda.synchronize()
    for k, b in bufs.items():
        assert b.intact(), k
        assert b.written(), k
    # the same through packed rows (qs_policy_pack_rollout): the pack stays inside its n * 128 bytes, and the statistics computed by
    # the library (adv_stats = NULL) and the gather through the rows leave the per-sample block unchanged up to the statistics
    packed = Guarded(n, 32, f32)
    _capi.check(L.qs_policy_pack_rollout(p(obs), p(act.view), p(ret), p(oldv), p(oldnlp), n, packed.ptr, _stream()))
    torch.cuda.synchronize()
    assert packed.intact() and packed.written() and torch.equal(packed.view[:, :12], obs) and torch.equal(packed.view[:, 16], ret)
    ps_arrays = bufs["per_sample"].view.clone()
    g.packed_rows = packed.view.data_ptr()
    _capi.check(L.qs_policy_ppo_grad(fused._h, C.byref(g), _stream()))
    torch.cuda.synchronize()
    assert torch.equal(ps_arrays, bufs["per_sample"].view)
    for k, b in bufs.items():
        assert b.intact(), k
    assert bool(torch.isfinite(bufs["per_sample"].view).all()) and bool(torch.isfinite(bufs["g2"].view).all())
    # every byte of every image is written (rows past the minibatch as zeros): nothing of the sentinel survives inside
    for k in ("a1", "h", "hpv", "dz2", "dz1", "dout"):
        v = bufs[k].view.view(torch.bfloat16).float()
        assert bool(torch.isfinite(v).all()), k
    last_rows = bufs["hpv"].view.view(torch.bfloat16).view(tiles, 32, 16, 8, 8)[-1].permute(1, 2, 0, 3).reshape(128, 256)[B - (tiles - 1) * 128:]
    assert float(last_rows.float().abs().max()) == 0.0 if last_rows.numel() else True


@pytest.mark.parametrize("env_id,dtype", [("docking-v2", torch.float32), ("docking-v0", torch.float64), ("hovering-v0", torch.float32),
                                          ("hovering-v0", torch.float64)])
def test_step_n_expert_peek_stay_in_bounds(env_id, dtype):
    n, T = 1000 + 77, 9
    env = qs.make_vec(env_id, n, dtype=dtype, seed=2)
    env.reset()
    od = env.obs_dim
    acts = torch.rand((T, n, 4), device="cuda", dtype=torch.float32).to(dtype)
    obs, rew, done = Guarded(T * n, od, dtype), Guarded(T * n, 0, dtype), Guarded(T * n, 0, torch.uint8)
    _capi.check(env.lib.qs_step_n(env._h, T, C.c_void_p(acts.data_ptr()), obs.ptr, rew.ptr, done.ptr, _stream()))
    torch.cuda.synchronize()
    assert obs.intact() and rew.intact() and done.intact() and obs.written() and rew.written()
    assert bool((done.view <= 1).all())
    ids = torch.tensor([0, n - 1, 128, 5], dtype=torch.int64, device="cuda")
    ch, tg, tt = Guarded(4, 13, torch.float64), Guarded(4, 13, torch.float64), Guarded(4, 0, torch.int32)
    _capi.check(env.lib.qs_peek_state(env._h, C.c_void_p(ids.data_ptr()), 4, ch.ptr, tg.ptr if env_id != "hovering-v0" else None, tt.ptr, _stream()))
    torch.cuda.synchronize()
    assert ch.intact() and tg.intact() and tt.intact() and ch.written()
    if env_id != "hovering-v0":
        mem, a_out, o_io = Guarded(n, 8, dtype), Guarded(n, 4, dtype), Guarded(n, 12, dtype)
        mem.view.zero_()
        mem.view[:, 4] = 1
        o_io.view.copy_(env.reset())
        _capi.check(env.lib.qs_expert_action(env._h, mem.ptr, 1, a_out.ptr, o_io.ptr, _stream()))
        torch.cuda.synchronize()
        assert mem.intact() and a_out.intact() and o_io.intact() and a_out.written()
    env.close()


def test_rollout_stays_in_bounds():
    n, T = 1000 + 5, 6
    env = qs.make_vec("docking-v2", n, seed=3, randomize="episode", want_info=False)
    pol = ppo.MlpPolicy().cuda()
    fused = FusedPolicy(pol)
    f32 = torch.float32
    obs, act, val = Guarded((T + 1) * n, 12, f32), Guarded(T * n, 4, f32), Guarded((T + 1) * n, 0, f32)
    nlp, rew, done, clip = Guarded(T * n, 0, f32), Guarded(T * n, 0, f32), Guarded((T + 1) * n, 0, torch.uint8), Guarded(n, 4, f32)
    obs.view[:n].copy_(env.reset())
    done.view[:n].zero_()
    for slices in (1, 2, 0):  # 0: the fused step kernel (policy forward + env step in one launch)
        b = _capi.RolloutBuffers(*[x.view.data_ptr() for x in (obs, act, val, nlp, rew, done, clip)])
        _capi.check(env.lib.qs_rollout(env._h, fused._h, T, C.byref(b), 0, 0, slices, _stream()))
        torch.cuda.synchronize()
        for name, x in dict(obs=obs, act=act, val=val, nlp=nlp, rew=rew, done=done, clip=clip).items():
            assert x.intact(), (slices, name)
        assert not bool((val.view == val.sentinel).any()) and not bool((rew.view == rew.sentinel).any())
    env.close()
