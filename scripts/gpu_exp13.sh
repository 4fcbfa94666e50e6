mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_umma.py -k "spade" > gpurun_out/t_spade.log 2>&1; echo "spade rc=$?"; tail -n 12 gpurun_out/t_spade.log | cut -c1-300
timeout 900 $PT tests/test_gpu_generator.py > gpurun_out/t_gen.log 2>&1; echo "gen rc=$?"; tail -n 3 gpurun_out/t_gen.log | cut -c1-300; grep -n "relL2" gpurun_out/t_gen.log | tail -n 10
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
head -n 1 gpurun_out/prof_convs_b64.log; grep -E "Cout1024|Cout512 64x64|Cout256 128x128|instnorm" gpurun_out/prof_convs_b64.log | head -n 20 | cut -c1-140
