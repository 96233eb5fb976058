mkdir -p gpurun_out/r2ak
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 3 -c 1 -o gpurun_out/r2ak/ppo_grad_kernel -f python tools/update_profile.py > gpurun_out/r2ak/ncu_grad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_env_step_kernel -s 20 -c 1 -o gpurun_out/r2ak/fused_step_kernel -f python tools/rollout_bench.py --fused --n-steps 16 --slices 0 > gpurun_out/r2ak/ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_fwd_ws_kernel_lean -s 10 -c 1 -o gpurun_out/r2ak/policy_ws_lean -f python tools/rollout_bench.py --fused --n-steps 16 --slices 2 > gpurun_out/r2ak/ncu_lean.log 2>&1
ls -la gpurun_out/r2ak
