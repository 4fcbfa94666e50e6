mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_umma_kernel -s 12 -c 2 -o gpurun_out/conv_umma_full python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_conv.log 2>&1; echo "ncu-conv rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rasterize_kernel -c 1 -o gpurun_out/rast_full python scripts/bench_rasterizer.py 256 > gpurun_out/ncu_rast.log 2>&1; echo "ncu-rast rc=$?"
tail -n 2 gpurun_out/bench.log gpurun_out/bench_ref.log
