mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_umma_kernel -s 422 -c 211 --csv --log-file gpurun_out/conv_traffic.csv python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 600 python scripts/bench_rasterizer.py 8192 > gpurun_out/bench_rast.log 2>&1
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
tail -n 1 gpurun_out/bench.log | cut -c1-400; tail -n 2 gpurun_out/bench_rast.log
