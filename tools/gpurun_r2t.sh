mkdir -p gpurun_out/r2t
cd $GRAFT_REPO_ROOT
for v in $VARIANTS; do
  if [ $v = default ]; then unset QUADSIM_B200_LIB; else export QUADSIM_B200_LIB=variants/$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_fused_policy.py -m gpu -q -x -k "forward_matches or shipped" 2>&1 | tail -1
  timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', {k: d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s')})"
done 2>&1 | tee gpurun_out/r2t/$LOGNAME.log
if [ -n "$TRACE" ]; then for v in $TRACE; do QUADSIM_B200_LIB=variants/$v.so timeout 300 python tools/ws_trace.py > gpurun_out/r2t/$v.json 2> gpurun_out/r2t/$v.err; done; fi
