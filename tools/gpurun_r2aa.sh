mkdir -p gpurun_out/r2aa
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_ppo_grad.py tests/test_gpu_guards.py tests/test_gpu_fused_policy.py -m gpu -q -x 2>&1 | tail -3)
QUADSIM_B200_LIB=variants/train_trace.so timeout 300 python tools/train_trace.py > gpurun_out/r2aa/train_trace.json 2> gpurun_out/r2aa/train_trace.err
tail -3 gpurun_out/r2aa/train_trace.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2aa/train_trace.json"))
print("period", d["period_clk"])
for k, v in d["phases"]: print("  %-50s %8.1f" % (k, v))
PY
timeout 300 python tools/update_profile.py 2>&1 | grep -v "^void at\|Memset\|Memcpy\|warn" | tail -9 | cut -c1-160
