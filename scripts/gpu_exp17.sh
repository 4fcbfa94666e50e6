mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py tests/test_gpu_halo.py > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 14 gpurun_out/t_all.log | cut -c1-300
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
head -n 16 gpurun_out/prof_convs_b64.log | cut -c1-130
