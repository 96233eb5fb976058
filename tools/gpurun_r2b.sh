mkdir -p gpurun_out/r2b
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150) > gpurun_out/r2b/pytest.log
echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b/bench_20.json 2> gpurun_out/r2b/bench_20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b/bench_ref.json 2> gpurun_out/r2b/bench_ref.err
QUADSIM_B200_NO_EXT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-callers > gpurun_out/r2b/bench_20_ctypes.json 2> gpurun_out/r2b/bench_20_ctypes.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b/launches_headline.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-callers --no-extras > gpurun_out/r2b/ncu_launch.log 2>&1
tail -5 gpurun_out/r2b/pytest.log
