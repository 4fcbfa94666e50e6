mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_umma_kernel -s 262 -c 3 -o gpurun_out/conv_umma_512_full -f python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_conv512.log 2>&1; echo "ncu conv512 rc=$?"
