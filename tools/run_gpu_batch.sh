cd $GRAFT_REPO_ROOT
timeout 300 python tools/opt_probe.py 2>&1 | grep stage
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_t11.log
cat gpurun_out/r2_t11.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run A=1
run MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=80 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=96 MSR3D_FFN_SLICE=64
} > gpurun_out/r2_sweep7.log 2>&1
cat gpurun_out/r2_sweep7.log
MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64 timeout 300 python tools/step_timeline.py 2>&1 | cut -c1-110 > gpurun_out/r2_step_tl3_88.log
