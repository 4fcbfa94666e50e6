mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_rasterizer.py > gpurun_out/t_rast.log 2>&1; echo "rast rc=$?"
timeout 900 $PT tests/test_gpu_ops.py -k "block_extract or local_attn" > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 600 python scripts/bench_rasterizer.py 8192 > gpurun_out/bench_rast.log 2>&1; echo "brast rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_b8.csv python scripts/profile_convs.py 8 bf16 > gpurun_out/ncu_b8.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/t_rast.log gpurun_out/t_ops.log gpurun_out/bench_rast.log; head -45 gpurun_out/prof_convs_b64.log
