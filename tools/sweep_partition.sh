run() { r=$(env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
run MSR3D_CHAIN_SMS=96 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=80 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=72 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=64 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=72 MSR3D_FFN_SLICE=64 MSR3D_TUNE=4=2,5=2
run MSR3D_CHAIN_SMS=64 MSR3D_FFN_SLICE=64 MSR3D_TUNE=4=2,5=2
run MSR3D_CHAIN_SMS=64 MSR3D_FFN_SLICE=64 MSR3D_TUNE=4=2
run MSR3D_CHAIN_SMS=88 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=104 MSR3D_FFN_SLICE=64
run MSR3D_CHAIN_SMS=112 MSR3D_FFN_SLICE=32
