# 8-GPU records of the final build: BASELINE config 5 (64 Mi envs, PPO2 rollout-in-the-loop) and the bench line at N=8
mkdir -p gpurun_out/r2ad
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29621 tools/config5_bench.py --iters 3 2>&1 | grep '^{' | tail -1) > gpurun_out/r2ad/config5_8gpu.json
cut -c1-600 gpurun_out/r2ad/config5_8gpu.json
(timeout 600 $TR --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1) > gpurun_out/r2ad/bench_n8.json
cut -c1-300 gpurun_out/r2ad/bench_n8.json
(timeout 600 $TR --master-port 29623 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>&1 | grep '^{' | tail -1) > gpurun_out/r2ad/bench_ref_n8.json
cut -c1-300 gpurun_out/r2ad/bench_ref_n8.json
