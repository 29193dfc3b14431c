cd $GRAFT_REPO_ROOT
run() { r=$(env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --fast 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run A=1
run MSR3D_TUNE=10=1
run MSR3D_TUNE=10=1 MSR3D_CHAIN_SMS=80
run MSR3D_TUNE=10=1 MSR3D_CHAIN_SMS=72
run A=2
run MSR3D_TUNE=10=1
run MSR3D_TUNE=10=1,12=3
run MSR3D_TUNE=10=1 MSR3D_FFN_SLICE=32
run MSR3D_TUNE=10=1 MSR3D_CHAIN_SMS=0
} > gpurun_out/r2_sweep12.log 2>&1
cat gpurun_out/r2_sweep12.log
