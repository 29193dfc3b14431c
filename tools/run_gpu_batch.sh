cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { r=$(env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --fast 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value'])); print({k[:28]:round(v['us_per_step'],1) for k,v in d['kernels'].items() if 'attn' in k})"); echo "$*: $r"; }
{
run MSR3D_ATTN_BWD_TC=0
run MSR3D_ATTN_BWD_TC=1
run MSR3D_ATTN_BWD_TC=1 MSR3D_CHAIN_SMS=72
run MSR3D_ATTN_BWD_TC=1 MSR3D_CHAIN_SMS=64
} > gpurun_out/r2_sweep14.log 2>&1
cat gpurun_out/r2_sweep14.log
timeout 300 python tools/chain_profile.py bwd 2>&1 | cut -c1-100 | grep -v "gemm3x_kernel<128>\|colsum" | tail -32 > gpurun_out/r2_chain_bwd2.log
cat gpurun_out/r2_chain_bwd2.log
