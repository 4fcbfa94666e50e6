mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py tests/test_gpu_generator.py > gpurun_out/t_ops.log 2>&1; echo "ops+gen rc=$?"; tail -n 4 gpurun_out/t_ops.log | cut -c1-300
grep -n "relL2" gpurun_out/t_ops.log | tail -n 10
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
head -n 1 gpurun_out/prof_convs_b64.log; grep -E "Cin64 Cout128 |Cin128 Cout128|Cin16|Cin8|seg_" gpurun_out/prof_convs_b64.log | head -n 20 | cut -c1-140
