mkdir -p gpurun_out/r2a
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120) > gpurun_out/r2a/pytest.log
echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_20.json 2> gpurun_out/r2a/bench_20.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a/bench_default.json 2> gpurun_out/r2a/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dock_step_kernel -s 12 -c 1 -o gpurun_out/r2a/dock_f32 -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-callers --no-extras > gpurun_out/r2a/ncu_f32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dock_step_pipe_kernel -s 12 -c 1 -o gpurun_out/r2a/dock_f64 -f python bench.py --dtype f64 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-callers --no-extras > gpurun_out/r2a/ncu_f64.log 2>&1
ls -la gpurun_out/r2a
tail -5 gpurun_out/r2a/pytest.log
