mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py -k "norm or fold" > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 1200 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
tail -n 2 gpurun_out/t_ops.log gpurun_out/t_gen.log; head -n 14 gpurun_out/prof_convs_b64.log; grep instnorm gpurun_out/prof_convs_b64.log
