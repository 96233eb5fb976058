mkdir -p gpurun_out/r2n
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 600 $TR --master-port 29611 tools/dist_ppo_check.py 2>&1 | head -150) > gpurun_out/r2n/dist_ppo_check.log
grep -v "^  File\|^    " gpurun_out/r2n/dist_ppo_check.log | head -40 | cut -c1-300
