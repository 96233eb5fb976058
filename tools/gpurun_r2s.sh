mkdir -p gpurun_out/r2s
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_fused_policy.py tests/test_gpu_ppo.py tests/test_gpu_guards.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2s/pytest.log
tail -5 gpurun_out/r2s/pytest.log
timeout 120 python - <<'PY' 2>&1 | tail -5
import torch, time
from quadsim_b200 import ppo
from quadsim_b200.fused_policy import FusedPolicy
pol = ppo.MlpPolicy().cuda()
obs = torch.randn(1 << 20, 12, device="cuda")
for prec in ("bf16", "fp32"):
    f = FusedPolicy(pol, precision=prec)
    for _ in range(3): f.step(obs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f.step(obs)
    e1.record(); torch.cuda.synchronize()
    print(prec, "us per 2^20 envs:", round(e0.elapsed_time(e1) / 20 * 1000, 1))
PY
