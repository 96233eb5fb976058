mkdir -p gpurun_out/r2k
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120) > gpurun_out/r2k/pytest.log
tail -4 gpurun_out/r2k/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench_20.json 2> gpurun_out/r2k/bench_20.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2k/launches_with_callers.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2k/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_kernel -s 3 -c 1 -o gpurun_out/r2k/dw_kernel -f python tools/update_profile.py > gpurun_out/r2k/ncu_dw.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 3 -c 1 -o gpurun_out/r2k/ppo_grad_kernel -f python tools/update_profile.py > gpurun_out/r2k/ncu_grad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_env_step_kernel -s 20 -c 1 -o gpurun_out/r2k/fused_step_kernel -f python tools/rollout_bench.py --fused --n-steps 16 --slices 0 > gpurun_out/r2k/ncu_fused.log 2>&1
ls -la gpurun_out/r2k
