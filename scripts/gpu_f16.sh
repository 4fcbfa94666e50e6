mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 900 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dtype f16 > gpurun_out/bench_f16.log 2>&1; echo "bench rc=$?"
tail -n 3 gpurun_out/t_ops.log gpurun_out/t_umma.log gpurun_out/t_gen.log; grep -n "relL2" gpurun_out/t_gen.log | tail -n 24; tail -n 1 gpurun_out/bench_f16.log | cut -c1-700
