mkdir -p gpurun_out/r2j
cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gpu_ppo.py -m gpu -x -q -s -k "fused_step_kernel" 2>&1 | tail -8) > gpurun_out/r2j/pytest_fused.log
tail -5 gpurun_out/r2j/pytest_fused.log
for s in 0 2; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bulk state prefetch: slices', d['slices'], 'rollout_ms', round(d['rollout_ms'],3), 'env-steps/s %.3g' % d['rollout_env_steps_per_s'])" >> gpurun_out/r2j/fused_rollout.log; done
cat gpurun_out/r2j/fused_rollout.log
