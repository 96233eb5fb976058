mkdir -p gpurun_out/r2ae
cd $GRAFT_REPO_ROOT
{
echo "== run_sim_PID"; timeout 120 python examples/run_sim_PID.py 2>&1 | tail -4
echo "== run_moving_docking_ppo2"; timeout 300 python examples/run_moving_docking_ppo2.py --n 16384 --n-steps 16 --total-timesteps 2e6 --save-path gpurun_out/r2ae/logs 2>&1 | tail -4
ls gpurun_out/r2ae/logs 2>/dev/null | head -3
Z=$(ls gpurun_out/r2ae/logs/*.zip 2>/dev/null | head -1)
echo "== run_trained_moving_docking_ppo2 $Z"; [ -n "$Z" ] && timeout 120 python examples/run_trained_moving_docking_ppo2.py $Z docking-v2 2>&1 | tail -3
echo "== run_expert_record"; timeout 200 python examples/run_expert_record.py 8 8 gpurun_out/r2ae/expert 2>&1 | tail -3
echo "== gym_docking drop-in"; timeout 120 python - <<'PY' 2>&1 | tail -3
import gym_docking
env = gym_docking.make("gym_docking:docking-v2")
obs = env.reset()
tot = 0.0
for k in range(50):
    obs, r, done, info = env.step(env.action_space.sample())
    tot += r
    if done: break
print("gym_docking ok", k, round(float(tot), 4), sorted(info)[:4])
PY
} 2>&1 | tee gpurun_out/r2ae/examples.log
rm -rf gpurun_out/r2ae/logs gpurun_out/r2ae/expert*
