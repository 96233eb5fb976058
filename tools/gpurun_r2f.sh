mkdir -p gpurun_out/r2f
cd $GRAFT_REPO_ROOT
for v in ws_chunk64 ws_chunk128; do
  QUADSIM_B200_NO_EXT=1 QUADSIM_B200_LIB=$PWD/variants/$v.so timeout 200 python tools/rollout_bench.py --fused --n-steps 16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'policy_fwd_ms', round(d['policy_fwd_ms'],4), 'rollout_ms', round(d['rollout_ms'],3), 'env-steps/s %.3g' % d['rollout_env_steps_per_s'], 'update_ms', round(d['update_ms'],2))" >> gpurun_out/r2f/chunk.log
done
cat gpurun_out/r2f/chunk.log
QUADSIM_B200_NO_EXT=1 QUADSIM_B200_LIB=$PWD/variants/ws_chunk128.so timeout 300 python -m pytest tests/test_gpu_fused_policy.py -m gpu -x -q 2>&1 | tail -3
