mkdir -p gpurun_out/r2u
cd $GRAFT_REPO_ROOT
{
echo "== e2e host pipeline A/B (tools/e2e_sweep.py, 2^20 envs, 60 steps)"
for rep in 1 2; do
python tools/e2e_sweep.py 2>/dev/null | sed 's/^/4-stream round robin: /'
QUADSIM_B200_HOST_PIPE=2 python tools/e2e_sweep.py 2>/dev/null | sed 's/^/upload+compute stream, download stream: /'
done
QUADSIM_B200_HOST_PIPE=2 QUADSIM_B200_HOST_CHUNK0=16384 python tools/e2e_sweep.py 2>/dev/null | sed 's/^/2-stream, first chunk 16Ki: /'
QUADSIM_B200_HOST_PIPE=2 QUADSIM_B200_HOST_CHUNK_CAP=524288 python tools/e2e_sweep.py 2>/dev/null | sed 's/^/2-stream, cap 512Ki: /'
echo "== fp64 pipe kernel start skew (tools/kbench.py --dtype f64)"
python tools/kbench.py --dtype f64 --warmup 100 --steps 40 --reps 5 "" variants/pipeskew100.so variants/pipeskew200.so 2>/dev/null | cut -c1-200
python tools/kbench.py --dtype f64 --warmup 100 --steps 40 --reps 5 "" variants/pipeskew100.so variants/pipeskew200.so 2>/dev/null | cut -c1-200
} 2>&1 | tee gpurun_out/r2u/ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_fwd_ws_kernel -s 10 -c 1 -o gpurun_out/r2u/policy_ws -f python tools/rollout_bench.py --fused --n-steps 16 --slices 1 > gpurun_out/r2u/ncu_policy.log 2>&1
ls -la gpurun_out/r2u
