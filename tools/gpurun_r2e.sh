mkdir -p gpurun_out/r2e
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_ppo_grad.py tests/test_gpu_ppo.py -m gpu -x -q -s 2>&1 | tail -30) > gpurun_out/r2e/pytest_ppo.log
tail -8 gpurun_out/r2e/pytest_ppo.log
timeout 300 python tools/update_profile.py > gpurun_out/r2e/update_profile.log 2>&1; tail -14 gpurun_out/r2e/update_profile.log | cut -c1-150
