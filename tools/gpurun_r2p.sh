mkdir -p gpurun_out/r2p
cd $GRAFT_REPO_ROOT
for v in ws_trace_lock; do
QUADSIM_B200_LIB=variants/$v.so timeout 300 python tools/ws_trace.py > gpurun_out/r2p/$v.json 2> gpurun_out/r2p/$v.err
tail -3 gpurun_out/r2p/$v.err
done
