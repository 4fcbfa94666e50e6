mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 900 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rasterize_kernel -c 1 -o gpurun_out/rast_full python scripts/bench_rasterizer.py 256 > gpurun_out/ncu_rast.log 2>&1; echo "ncu rc=$?"
tail -4 gpurun_out/t_ops.log gpurun_out/t_umma.log gpurun_out/t_gen.log; head -40 gpurun_out/prof_convs_b64.log
