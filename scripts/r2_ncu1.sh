mkdir -p gpurun_out
for s in stem c128_64 convT128; do
  timeout 300 python scripts/one_conv.py $s 5
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:conv_umma_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2_ncu_$s python scripts/one_conv.py $s 3 > gpurun_out/r2_ncu_$s.log 2>&1; tail -2 gpurun_out/r2_ncu_$s.log
done
ls -la gpurun_out/*.ncu-rep
