PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py 2>&1 | tail -5
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r2_prof_tapskip.log 2>&1; head -n 40 gpurun_out/r2_prof_tapskip.log | cut -c1-150
