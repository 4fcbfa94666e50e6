mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_halo.py tests/test_gpu_ops.py > gpurun_out/t_umma.log 2>&1; echo "umma+halo+ops rc=$?"; tail -n 2 gpurun_out/t_umma.log
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
head -n 1 gpurun_out/prof_convs_b64.log; grep -E "conv2d" gpurun_out/prof_convs_b64.log | head -n 22
HOIG_UMMA_DEBUG=3 timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dbg3.log 2>&1
head -n 1 gpurun_out/prof_dbg3.log; grep -E "conv2d" gpurun_out/prof_dbg3.log | head -n 12
