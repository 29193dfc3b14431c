cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t12.log
cat gpurun_out/r2_t12.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run MSR3D_FFN_PRESPLIT=0 MSR3D_TUNE=12=2 MSR3D_FFN_SLICE=64
run MSR3D_FFN_PRESPLIT=1 MSR3D_TUNE=12=2 MSR3D_FFN_SLICE=64
run MSR3D_FFN_PRESPLIT=1 MSR3D_TUNE=12=2 MSR3D_FFN_SLICE=32
run MSR3D_FFN_PRESPLIT=1 MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64 MSR3D_TUNE=12=2
run MSR3D_FFN_PRESPLIT=1 MSR3D_CHAIN_SMS=80 MSR3D_FFN_SLICE=64 MSR3D_TUNE=12=2
} > gpurun_out/r2_sweep9.log 2>&1
cat gpurun_out/r2_sweep9.log
timeout 300 python tools/chain_profile.py fwd 2>&1 | cut -c1-100 | tail -14 > gpurun_out/r2_chain6.log
cat gpurun_out/r2_chain6.log
