mkdir -p gpurun_out/r2ai
cd $GRAFT_REPO_ROOT
run() { timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 slices $1', {k: round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('rollout_ms','rollout_env_steps_per_s')})"; }
{
python tools/kbench.py --warmup 100 --steps 200 --reps 3 "" variants/env_cg.so variants/env_cs.so 2>/dev/null | cut -c1-200
for v in default env_cg env_cs; do
  if [ $v = default ]; then unset QUADSIM_B200_LIB; else export QUADSIM_B200_LIB=variants/$v.so; fi
  run 2 $v; run 0 $v
done
} 2>&1 | tee gpurun_out/r2ai/stream.log
