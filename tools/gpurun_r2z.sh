mkdir -p gpurun_out/r2z
cd $GRAFT_REPO_ROOT
run() { timeout 200 python tools/rollout_bench.py --fused --n-steps 16 --slices $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 slices $1', {k: round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('policy_fwd_ms','rollout_ms','rollout_env_steps_per_s')})"; }
{
for s in 1 2 3 4; do run $s default; done
for r in 96 80; do export QUADSIM_B200_POLICY_REGS=$r; for s in 2 3; do run $s regs$r; done; done
unset QUADSIM_B200_POLICY_REGS
export QUADSIM_B200_POLICY_PRIO=0; run 2 prio0; unset QUADSIM_B200_POLICY_PRIO
for c in 132 116; do export QUADSIM_B200_POLICY_CTAS=$c; for s in 2 4; do run $s ctas$c; done; done
} 2>&1 | tee gpurun_out/r2z/rollout_sweep.log
