mkdir -p gpurun_out/r2ah
cd $GRAFT_REPO_ROOT
run() { timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 slices $1', {k: round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s')})"; }
{
for v in default fu2sets; do
  if [ $v = default ]; then unset QUADSIM_B200_LIB; else export QUADSIM_B200_LIB=variants/$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_ppo.py -m gpu -q -x -k fused_step_kernel 2>&1 | tail -1
  run 0 $v; run 0 $v
done
} 2>&1 | tee gpurun_out/r2ah/fused_sets.log
