mkdir -p gpurun_out/r2w
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_ppo_grad.py tests/test_gpu_ppo.py tests/test_gpu_guards.py -m gpu -q -x 2>&1 | tail -40) > gpurun_out/r2w/pytest.log
tail -3 gpurun_out/r2w/pytest.log
for i in 1 2; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s','update_ms_per_epoch')})"; done | tee gpurun_out/r2w/bench.log
timeout 300 python tools/update_profile.py 2>&1 | tail -30 > gpurun_out/r2w/update_profile.log
tail -25 gpurun_out/r2w/update_profile.log
