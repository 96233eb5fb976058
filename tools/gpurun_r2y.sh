mkdir -p gpurun_out/r2y
cd $GRAFT_REPO_ROOT
export QUADSIM_B200_LIB=variants/pipe_nosync.so
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "f64 or fp64 or float64 or oracle" 2>&1 | tail -4)
unset QUADSIM_B200_LIB
for i in 1 2; do python tools/kbench.py --dtype f64 --warmup 100 --steps 40 --reps 5 "" variants/pipe_nosync.so 2>/dev/null | cut -c1-230; done | tee gpurun_out/r2y/nosync.log
python tools/kbench.py --dtype f64 --warmup 300 --steps 2000 --reps 1 "" variants/pipe_nosync.so 2>/dev/null | cut -c1-230 | tee -a gpurun_out/r2y/nosync.log
