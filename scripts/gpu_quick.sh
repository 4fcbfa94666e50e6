mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 900 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 2 gpurun_out/t_ops.log gpurun_out/t_umma.log gpurun_out/t_gen.log; head -n 24 gpurun_out/prof_convs_b64.log; tail -n 1 gpurun_out/bench.log | cut -c1-1500
timeout 900 python -m pytest -m gpu -q -s -p no:cacheprovider tests/test_gpu_boundaries.py > gpurun_out/t_bound.log 2>&1; echo "bound rc=$?"; tail -n 15 gpurun_out/t_bound.log
