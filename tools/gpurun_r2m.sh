# 2-GPU validation of the round-2 build: PPO2 caller over NCCL, both bench arms under torchrun, smoke()
mkdir -p gpurun_out/r2m
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29611 tools/dist_ppo_check.py 2>&1 | tail -40) > gpurun_out/r2m/dist_ppo_check.log
tail -3 gpurun_out/r2m/dist_ppo_check.log
(timeout 900 $TR --master-port 29612 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>&1 | tail -20) > gpurun_out/r2m/bench_ref_n2.log
tail -2 gpurun_out/r2m/bench_ref_n2.log | cut -c1-600
(timeout 900 $TR --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -20) > gpurun_out/r2m/bench_n2.log
tail -2 gpurun_out/r2m/bench_n2.log | cut -c1-1500
(timeout 600 python __graft_entry__.py smoke 2>&1 | tail -20) > gpurun_out/r2m/smoke.log
tail -3 gpurun_out/r2m/smoke.log
