mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:attn_combine_tc --launch-skip 20 --launch-count 1 -f -o gpurun_out/r2_ncu_attn_tc python scripts/profile_convs.py 8 f16 > gpurun_out/r2_ncu_attn_tc.log 2>&1; tail -2 gpurun_out/r2_ncu_attn_tc.log
