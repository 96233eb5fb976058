mkdir -p gpurun_out/r2l
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150) > gpurun_out/r2l/pytest.log
tail -6 gpurun_out/r2l/pytest.log
