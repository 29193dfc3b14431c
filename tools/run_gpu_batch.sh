cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_t14.log
cat gpurun_out/r2_t14.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --fast 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value'])); print({k:round(v['us_per_step'],1) for k,v in d['kernels'].items() if 'linear_t' in k})"); echo "$*: $r"; }
{
run MSR3D_SA3_CHAIN=0
run MSR3D_SA3_CHAIN=1
run MSR3D_SA3_CHAIN=1 MSR3D_CHAIN_SMS=88
run MSR3D_SA3_CHAIN=1 MSR3D_CHAIN_SMS=96
} > gpurun_out/r2_sweep13.log 2>&1
cat gpurun_out/r2_sweep13.log
