cd $GRAFT_REPO_ROOT
run() { r=$(env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run MSR3D_TUNE=12=2
run MSR3D_TUNE=12=4
run MSR3D_TUNE=12=8
run MSR3D_TUNE=12=2 MSR3D_FFN_SLICE=64
run MSR3D_TUNE=12=4 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64 MSR3D_TUNE=12=2
run MSR3D_CHAIN_SMS=104 MSR3D_FFN_SLICE=64 MSR3D_TUNE=12=2
} > gpurun_out/r2_sweep8.log 2>&1
cat gpurun_out/r2_sweep8.log
