mkdir -p gpurun_out/r2aj
cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6 7 8; do
  timeout 600 python -m pytest tests/test_gpu_guards.py tests/test_gpu_parity.py "tests/test_gpu_parity_fullsize.py::test_docking_v2_full_size_vs_oracle" -m gpu -q -x 2>&1 > gpurun_out/r2aj/combo_$i.log
  tail -1 gpurun_out/r2aj/combo_$i.log
  grep -n "AssertionError" -B14 gpurun_out/r2aj/combo_$i.log | cut -c1-260 | head -40
done
