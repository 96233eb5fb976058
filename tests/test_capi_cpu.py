"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol include/quadsim_b200.h
declares, reports errors without throwing, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from quadsim_b200 import _capi
from qs_testutil import ROOT


def test_header_and_binding_agree():
    hdr = open(os.path.join(ROOT, "include", "quadsim_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qs_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_capi.SYMBOLS), (declared, sorted(_capi.SYMBOLS))
    assert "#define QS_ABI_VERSION %d" % _capi.ABI_VERSION in hdr
    for name, val in (("QS_HOVER", _capi.HOVER), ("QS_DOCK_V0", _capi.DOCK_V0), ("QS_DOCK_V2", _capi.DOCK_V2),
                      ("QS_F32", _capi.F32), ("QS_F64", _capi.F64), ("QS_RAND_EPISODE", _capi.RAND_EPISODE),
                      ("QS_RAND_ONCE", _capi.RAND_ONCE), ("QS_FN_STATE2REL", _capi.FN_STATE2REL), ("QS_FN_PHILOX", _capi.FN_PHILOX)):
        assert re.search(r"\b%s = %d\b" % (name, val), hdr), name


def test_library_exports_every_declared_symbol():
    L = _capi.load()
    for sym in _capi.SYMBOLS:
        assert hasattr(L, sym), sym
    assert L.qs_abi_version() == _capi.ABI_VERSION


def test_config_struct_layout_matches_header():
    # int32 int32 int64 uint64 uint64 int32 int32 double int32 int32 = 56 bytes, no implicit padding
    assert C.sizeof(_capi.Config) == 56
    assert _capi.Config.n_envs.offset == 8 and _capi.Config.half_width.offset == 40 and _capi.Config.auto_reset.offset == 48
    assert C.sizeof(_capi.StateArrays) == 9 * C.sizeof(C.c_void_p)


def test_invalid_arguments_return_status_not_crash():
    L = _capi.load()
    h = C.c_void_p()
    assert L.qs_create(C.byref(h), None) == -1 and b"null" in L.qs_last_error()
    bad = _capi.Config(7, 0, 16, 0, 0, 0, 0, 0.5, 1, 0)
    assert L.qs_create(C.byref(h), C.byref(bad)) == -1 and b"env_kind" in L.qs_last_error()
    bad = _capi.Config(2, 0, 0, 0, 0, 0, 0, 0.5, 1, 0)
    assert L.qs_create(C.byref(h), C.byref(bad)) == -1 and b"n_envs" in L.qs_last_error()
    assert L.qs_step(None, None, None, None, None, None) == -1
    assert L.qs_reset(None, None, None) == -1
    assert L.qs_get_stats(None, None, None) == -1
    assert L.qs_destroy(None) == 0
    assert L.qs_fn_apply(99, 0, 4, None, None, None, None, None, None) == -1
    # the callers' entry points refuse bad handles the same way
    assert L.qs_step_range(None, 0, 0, None, None, None, None, None) == -1
    assert L.qs_rollout(None, None, 4, None, 0, 0, 1, None) == -1 and b"qs_rollout" in L.qs_last_error()
    assert L.qs_policy_step(None, 4, None, 0, 0, 0, 0, None, None, None, None, None) == -1
    assert L.qs_policy_set_weights(None, None, None) == -1
    assert L.qs_policy_ppo_grad(None, None, None) == -1
    assert L.qs_policy_destroy(None) == 0
    assert L.qs_expert_action(None, None, 1, None, None, None) == -1
    assert L.qs_peek_state(None, None, 1, None, None, None, None) == -1
    assert L.qs_gae(0, 4, None, None, None, None, None, 0.99, 0.95, None, None, None) == -1
    assert L.qs_env_dtype(None) == -1 and L.qs_env_id0(None) == 0
    assert C.sizeof(_capi.RolloutBuffers) == 7 * C.sizeof(C.c_void_p)
    assert C.sizeof(_capi.PPOGrad) == 6 * 8 + 8 + 8 + 3 * 4 + 4 + 6 * 8  # 12-byte float block padded to the next pointer


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import quadsim_b200 as qs
    L = _capi.load()
    h = C.c_void_p()
    cfg = _capi.Config(2, 0, 16, 0, 0, 0, 0, 0.5, 1, 0)
    assert L.qs_create(C.byref(h), C.byref(cfg)) == -3  # QS_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.qs_last_error()
    with pytest.raises(qs.QuadSimError):
        qs.make_vec("docking-v2", 8)
    with pytest.raises(qs.QuadSimError):
        qs.make("docking-v0")
    from quadsim_b200 import functional as F
    with pytest.raises(qs.QuadSimError):
        F.quat2euler(torch.zeros(4, 4))
    hp = C.c_void_p()
    assert L.qs_policy_create(C.byref(hp), 0) == -3 and b"no CPU fallback" in L.qs_last_error()
    from quadsim_b200 import ppo
    from quadsim_b200.fused_policy import FusedPolicy
    with pytest.raises(qs.QuadSimError):
        FusedPolicy(ppo.MlpPolicy())


def test_product_never_imports_the_oracle():
    """The oracle and the host harness are test infrastructure: nothing in the package may import, link or load them."""
    pkg = os.path.join(ROOT, "quadsim_b200")
    banned = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|quadsim_oracle|oracle/|oracle\.py|qs_host_harness|import\s+host_harness")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not banned.search(src), os.path.join(dirpath, f)


def test_integration_doc_names_every_entry_point():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in _capi.SYMBOLS if s not in doc]
    assert not missing, missing
