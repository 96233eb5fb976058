mkdir -p gpurun_out/r2af
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_ppo.py -m gpu -q -x -k "gae or runner or in_place" 2>&1 | tail -2
timeout 120 python - <<'PY' 2>&1 | tail -3
import torch, ctypes as C
from quadsim_b200 import _capi
L = _capi.load()
n, T = 1 << 20, 16
rew = torch.randn(T, n, device="cuda"); val = torch.randn(T, n, device="cuda"); dones = (torch.rand(T, n, device="cuda") < 0.02).to(torch.uint8)
lv = torch.randn(n, device="cuda"); ld = torch.zeros(n, dtype=torch.uint8, device="cuda")
adv = torch.empty(T, n, device="cuda"); ret = torch.empty(T, n, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
def run():
    _capi.check(L.qs_gae(T, n, p(rew), p(val), p(dones), p(lv), p(ld), 0.99, 0.95, p(adv), p(ret), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(5): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
print("qs_gae 2^20 x 16: %.1f us" % (e0.elapsed_time(e1) / 50 * 1000))
PY
for i in 1 2; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('rollout_ms','rollout_env_steps_per_s')})"; done
