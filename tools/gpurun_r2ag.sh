mkdir -p gpurun_out/r2ag
cd $GRAFT_REPO_ROOT
run() { timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 slices $1', {k: round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s')})"; }
{
timeout 600 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_policy.py -m gpu -q -x 2>&1 | tail -1
run 2 default; run 2 default; run 1 default; run 0 default
QUADSIM_B200_POLICY_LEAN=0 run 2 lean0
export QUADSIM_B200_LIB=variants/fu_c32.so
timeout 300 python -m pytest tests/test_gpu_ppo.py -m gpu -q -x -k fused_step_kernel 2>&1 | tail -1
run 0 fu_c32
} 2>&1 | tee gpurun_out/r2ag/lean.log
