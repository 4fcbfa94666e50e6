mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_rasterizer.py > gpurun_out/t_rast.log 2>&1; echo "rast rc=$?"
timeout 900 $PT tests/test_gpu_ops.py > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 900 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 600 python scripts/bench_rasterizer.py 8192 > gpurun_out/bench_rast.log 2>&1; echo "brast rc=$?"
for f in t_rast t_ops t_umma t_gen bench_rast; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done; head -n 42 gpurun_out/prof_convs_b64.log
