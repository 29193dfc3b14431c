cd $GRAFT_REPO_ROOT
run() { r=$(env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run MSR3D_ATTN_BLOCK=0
run MSR3D_ATTN_BLOCK=1
run MSR3D_ATTN_BLOCK=1 MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64
run MSR3D_ATTN_BLOCK=1 MSR3D_CHAIN_SMS=72 MSR3D_FFN_SLICE=64
run MSR3D_ATTN_BLOCK=1 MSR3D_CHAIN_SMS=64 MSR3D_FFN_SLICE=64
run MSR3D_ATTN_BLOCK=1 MSR3D_CHAIN_SMS=96 MSR3D_FFN_SLICE=64
} > gpurun_out/r2_sweep5.log 2>&1
cat gpurun_out/r2_sweep5.log
timeout 300 python tools/chain_profile.py fwd 2>&1 | cut -c1-100 > gpurun_out/r2_chain5.log
