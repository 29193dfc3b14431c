cd $GRAFT_REPO_ROOT
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/step_launches.py > gpurun_out/r2_launches.log 2>&1
wc -l gpurun_out/r2_launches.csv
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"sa_mlp_kernel|gemm3x_kernel|attn_block|ffn_pair|linear_tma|attn_bwd" -c 40 -o /tmp/r2_prof -f python tools/step_launches.py > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw.csv 2>/dev/null
ncu -i /tmp/r2_prof.ncu-rep --page details --csv --kernel-name regex:"attn_block" 2>/dev/null | head -400 > gpurun_out/r2_ncu_attn_block_details.csv
wc -c gpurun_out/r2_ncu_raw.csv
