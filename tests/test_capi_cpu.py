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


def test_config_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors against the C compiler's own layout of include/quadsim_b200.h (sizeof / offsetof of every struct
    that crosses the boundary), so a field added on one side only cannot go unnoticed."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text("""
#include <stdio.h>
#include <stddef.h>
#include "quadsim_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(qs_config_t), offsetof(qs_config_t, n_envs), offsetof(qs_config_t, half_width),
         offsetof(qs_config_t, auto_reset), offsetof(qs_config_t, use_params), offsetof(qs_config_t, params), sizeof(qs_env_params_t));
  printf("%zu %zu %zu %zu\\n", offsetof(qs_env_params_t, inertia), offsetof(qs_env_params_t, kp_vel), offsetof(qs_env_params_t, dock_port),
         offsetof(qs_env_params_t, max_steps));
  printf("%zu %zu %zu %zu %zu\\n", sizeof(qs_state_arrays_t), sizeof(qs_step_extras_t), sizeof(qs_rollout_t), sizeof(qs_ppo_grad_t),
         sizeof(qs_policy_weights_t));
  printf("%zu\\n", sizeof(qs_adam_tensor_t));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    rows = [[int(x) for x in line.split()] for line in subprocess.check_output([str(exe)]).decode().splitlines()]
    Cfg, Prm = _capi.Config, _capi.EnvParams
    assert rows[0] == [C.sizeof(Cfg), Cfg.n_envs.offset, Cfg.half_width.offset, Cfg.auto_reset.offset, Cfg.use_params.offset,
                       Cfg.params.offset, C.sizeof(Prm)]
    assert rows[1] == [Prm.inertia.offset, Prm.kp_vel.offset, Prm.dock_port.offset, Prm.max_steps.offset]
    assert rows[2] == [C.sizeof(_capi.StateArrays), C.sizeof(_capi.StepExtras), C.sizeof(_capi.RolloutBuffers), C.sizeof(_capi.PPOGrad),
                       C.sizeof(_capi.PolicyWeights)]
    assert rows[3] == [C.sizeof(_capi.AdamTensor)]


def test_default_params_are_the_reference_literals():
    """qs_default_params == dynamics/quadrotor.py:10-63, controller/PIDController.py:10-44 and the envs' __init__ literals."""
    d0 = _capi.default_params(_capi.DOCK_V0).as_dict()
    d2 = _capi.default_params(_capi.DOCK_V2).as_dict()
    dh = _capi.default_params(_capi.HOVER).as_dict()
    assert (d2["dt"], d2["gravity"], d2["mass"], d2["arm_length"]) == (0.02, 9.81, 0.18, 0.086)
    assert d2["inertia"] == [0.00025, 0.000232, 0.0003738] and d2["motor_lambda"] == 1.5e-9 / 6.11e-8
    assert d2["att_limit_deg"] == [85, 85, 175]
    assert d2["kp_att"] == [-10, -10, -9.5] and d2["kd_att"] == [5.1, 5.1, 4.0]
    assert d2["kp_pos"] == [-1.0, -1.0, 50] and d2["kd_pos"] == [-1.65, -1.65, 8] and d2["kp_vel"] == [-0.7, -0.7, 1]
    assert d2["chaser_ini_pos"] == [8, -50, 5] and d2["target_ini_pos"] == [10, -50, 5] and d0["target_des_pos"] == [10, -50, 5]
    assert d2["target_des_vel"] == [0.2, 0, 0] and d0["target_des_vel"] == [0, 0, 0]
    assert (d0["range_limit"], d2["range_limit"], dh["range_limit"]) == (3, 10, 100)
    assert (d2["dock_port"], d2["dock_pos_tol"], d2["dock_vel_tol"], d2["dock_angle_deg"], d2["min_altitude"], d2["max_steps"]) == (0.1, 0.1, 0.1, 10, 0.1, 600)
    assert dh["chaser_ini_pos"] == [0, 0, 5]
    assert _capi.load().qs_default_params(9, None) == -1


def test_invalid_arguments_return_status_not_crash():
    L = _capi.load()
    h = C.c_void_p()
    assert L.qs_create(C.byref(h), None) == -1 and b"null" in L.qs_last_error()
    bad = _capi.Config(7, 0, 16, 0, 0, 0, 0, 0.5, 1, 0)
    assert L.qs_create(C.byref(h), C.byref(bad)) == -1 and b"env_kind" in L.qs_last_error()
    bad = _capi.Config(2, 0, 0, 0, 0, 0, 0, 0.5, 1, 0)
    assert L.qs_create(C.byref(h), C.byref(bad)) == -1 and b"n_envs" in L.qs_last_error()
    assert L.qs_step(None, None, None, None, None, None) == -1
    assert L.qs_step_ex(None, None, None, None, None, None, None) == -1
    assert L.qs_step_host(None, None, None, None, None, None) == -1
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
    assert L.qs_env_dtype(None) == -1 and L.qs_env_id0(None) == 0 and L.qs_env_device(None) == -1
    assert C.sizeof(_capi.RolloutBuffers) == 7 * C.sizeof(C.c_void_p)
    assert C.sizeof(_capi.PPOGrad) == 6 * 8 + 8 + 8 + 3 * 4 + 4 + 12 * 8  # 12-byte float block padded to the next pointer
    assert L.qs_adam_step(None, 0, 3e-4, 0.9, 0.999, 1e-5, 1, None, 0.5, None, None) == -1 and b"qs_adam_step" in L.qs_last_error()
    assert L.qs_permutation(0, 1, None, None) == -1


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
