cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -k "attached" 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/r2_bench_full2.json 2> gpurun_out/r2_bench_full2.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_full2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'])
print(json.dumps(d['secondary'])[:900])
print(d['roofline']['kernel'], d['roofline']['frac'], d['step_roofline']['frac_tensor'], d['gpu_baseline'].get('ratio_value_over_gpu_baseline'))
"
