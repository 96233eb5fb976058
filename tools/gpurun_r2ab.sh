mkdir -p gpurun_out/r2ab
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_ppo_grad.py tests/test_gpu_guards.py tests/test_gpu_ppo.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2ab/pytest.log; tail -3 gpurun_out/r2ab/pytest.log
timeout 300 python tools/update_profile.py 2>&1 | grep -v "^void at\|Memset\|Memcpy\|warn" | tail -10 | cut -c1-160 | tee gpurun_out/r2ab/update_profile.log
for i in 1 2; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('rollout_ms','update_ms_per_epoch')})"; done
timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices 2 --epochs 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('4 epochs', {k: d[k] for k in ('update_ms','update_ms_per_epoch')})"
