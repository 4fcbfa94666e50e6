mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rasterize_kernel -c 1 -o gpurun_out/rast2_full python scripts/bench_rasterizer.py 256 > gpurun_out/ncu_rast2.log 2>&1; echo "ncu rc=$?"
