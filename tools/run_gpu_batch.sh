cd $GRAFT_REPO_ROOT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/check_ddp_equiv.py > gpurun_out/r2_ddp_eq3.log 2>&1
echo "rc=$?"; grep "vs sequential\|equivalence\|Error\|assert" gpurun_out/r2_ddp_eq3.log | tail -5
run2() { r=$(env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 40 --warmup 5 --fast 2>gpurun_out/r2_ddp_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
run1() { r=$(env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --fast 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"); echo "$*: $r"; }
{
run1 N=1
run2 MSR3D_P2P_ALLREDUCE=1
run2 MSR3D_P2P_ALLREDUCE=0
} > gpurun_out/r2_ddp_sweep3.log 2>&1
cat gpurun_out/r2_ddp_sweep3.log
