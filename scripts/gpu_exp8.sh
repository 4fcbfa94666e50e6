mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_halo.py > gpurun_out/t_halo.log 2>&1; echo "halo rc=$?"; tail -n 3 gpurun_out/t_halo.log | cut -c1-300
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
head -n 1 gpurun_out/prof_convs_b64.log; grep -E "halo" gpurun_out/prof_convs_b64.log | head -n 3
