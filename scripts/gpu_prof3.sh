mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_umma_kernel -s 18 -c 1 -o gpurun_out/convT_full python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_convT.log 2>&1; echo "ncu rc=$?"
