mkdir -p gpurun_out/r2d
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_ppo_grad.py tests/test_gpu_ppo.py tests/test_gpu_fused_policy.py -m gpu -x -q -s 2>&1 | tail -60) > gpurun_out/r2d/pytest_ppo.log
tail -15 gpurun_out/r2d/pytest_ppo.log
timeout 300 python tools/rollout_bench.py --fused --n-steps 16 > gpurun_out/r2d/rollout_bench.json 2> gpurun_out/r2d/rollout_bench.err; cat gpurun_out/r2d/rollout_bench.json
timeout 300 python tools/update_profile.py > gpurun_out/r2d/update_profile.log 2>&1; tail -40 gpurun_out/r2d/update_profile.log
for c in imm bank; do QUADSIM_B200_CONSTS=$c timeout 300 python bench.py --dtype f64 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-callers --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp64 consts $c', d['ms_per_step'], d['roofline']['frac'])"; done >> gpurun_out/r2d/f64_consts.log; cat gpurun_out/r2d/f64_consts.log
