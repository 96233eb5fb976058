# final record of the round: full GPU suite, smoke, both bench arms (driver's invocation), launch list
mkdir -p gpurun_out/r2v
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -120) > gpurun_out/r2v/pytest.log
tail -3 gpurun_out/r2v/pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4) > gpurun_out/r2v/smoke.log
tail -2 gpurun_out/r2v/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2v/bench_ref.json 2> gpurun_out/r2v/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v/bench_20.json 2> gpurun_out/r2v/bench_20.err
timeout 900 python bench.py > gpurun_out/r2v/bench_default.json 2> gpurun_out/r2v/bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2v/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2v/ncu_launch.log 2>&1
ls -la gpurun_out/r2v
