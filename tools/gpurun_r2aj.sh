mkdir -p gpurun_out/r2aj
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2aj/full_$i.log
  tail -2 gpurun_out/r2aj/full_$i.log
  grep -n "AssertionError" -B12 gpurun_out/r2aj/full_$i.log | cut -c1-260 | head -60
done
