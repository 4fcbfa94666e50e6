mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_rasterizer.py > gpurun_out/t_rast.log 2>&1; echo "rast rc=$?"
for bp in 16384 8192 4096; do timeout 300 python scripts/bench_rasterizer.py 8192 $bp > gpurun_out/bench_rast_$bp.log 2>&1; echo "bp=$bp"; sed -n 2p gpurun_out/bench_rast_$bp.log | cut -c1-120; done
tail -n 3 gpurun_out/t_rast.log
