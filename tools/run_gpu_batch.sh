cd $GRAFT_REPO_ROOT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/check_ddp_equiv.py > gpurun_out/r2_ddp_eq.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/r2_ddp_eq.log
