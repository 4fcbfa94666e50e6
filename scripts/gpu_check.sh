mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_rasterizer.py > gpurun_out/t_rast.log 2>&1; echo "rast rc=$?"
timeout 900 $PT tests/test_gpu_ops.py > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 900 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/t_rast.log gpurun_out/t_ops.log gpurun_out/t_umma.log gpurun_out/t_gen.log gpurun_out/bench_bf16.log
