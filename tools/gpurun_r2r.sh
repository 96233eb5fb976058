mkdir -p gpurun_out/r2r
cd $GRAFT_REPO_ROOT
for v in $VARIANTS; do
  if [ $v = default ]; then unset QUADSIM_B200_LIB; else export QUADSIM_B200_LIB=variants/$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_ppo.py -m gpu -q -x -k "${K:-fused_step_kernel}" 2>&1 | tail -3
  for s in ${SLICES:-0 2}; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v slices $s', {k: d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s')})"; done
done 2>&1 | tee gpurun_out/r2r/$LOGNAME.log
