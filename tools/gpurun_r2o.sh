mkdir -p gpurun_out/r2o
cd $GRAFT_REPO_ROOT
for v in default variants/exp_l3steps4.so variants/exp_l3steps1.so; do
  if [ $v = default ]; then unset QUADSIM_B200_LIB; else export QUADSIM_B200_LIB=$v; fi
  for s in 1 2; do timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v slices $s', {k: d[k] for k in d if 'ms' in k or 'us' in k or 'per_s' in k})"; done
done 2>&1 | tee gpurun_out/r2o/l3steps.log
