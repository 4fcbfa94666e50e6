mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"; tail -n 14 gpurun_out/t_umma.log | cut -c1-300
for b in 1 0; do
HOIG_UMMA_HALO=$b timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_halo$b.log 2>&1
echo "== halo $b"; head -n 1 gpurun_out/prof_halo$b.log; grep -E "conv k3 s1 .*(256x256|128x128)" gpurun_out/prof_halo$b.log | cut -c1-130
done
