mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_halo.py > gpurun_out/t_halo.log 2>&1; echo "halo rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 4 gpurun_out/t_halo.log gpurun_out/t_gen.log; grep -n "relL2" gpurun_out/t_gen.log | tail -n 20; head -n 30 gpurun_out/prof_convs_b64.log; tail -n 1 gpurun_out/bench.log | cut -c1-600
