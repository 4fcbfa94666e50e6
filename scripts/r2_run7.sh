mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py > gpurun_out/r2_t_umma.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2_t_umma.log | cut -c1-300
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r2_prof_s123.log 2>&1; head -n 32 gpurun_out/r2_prof_s123.log | cut -c1-150
