mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py > gpurun_out/r2_t_umma.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2_t_umma.log | cut -c1-300
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r2_prof_vh1.log 2>&1; head -n 30 gpurun_out/r2_prof_vh1.log | cut -c1-150
HOIG_UMMA_VHALO=0 timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r2_prof_vh0.log 2>&1; grep "k7" gpurun_out/r2_prof_vh0.log | cut -c1-150
