mkdir -p gpurun_out/r2c
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "runtime_params or invalid_params or tapes" 2>&1 | tail -40) > gpurun_out/r2c/pytest_params.log
for regs in 128 96 80 72; do for s in 2 3 4; do
  QUADSIM_B200_POLICY_REGS=$regs timeout 120 python tools/rollout_bench.py --fused --n-steps 16 --slices $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('regs', $regs, 'slices', d['slices'], 'rollout_ms', round(d['rollout_ms'],3), 'env-steps/s %.3g' % d['rollout_env_steps_per_s'], 'fwd_ms', round(d.get('policy_forward_ms', 0), 4))" >> gpurun_out/r2c/sweep.log
done; done
for regs in 128 80 72; do QUADSIM_B200_POLICY_REGS=$regs timeout 120 python tools/overlap_probe.py 2>/dev/null | tail -3 >> gpurun_out/r2c/overlap.log; done
cat gpurun_out/r2c/sweep.log; cat gpurun_out/r2c/overlap.log; tail -5 gpurun_out/r2c/pytest_params.log
