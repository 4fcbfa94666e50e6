mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"; tail -n 12 gpurun_out/t_umma.log | cut -c1-300
for d in 1 0; do
HOIG_UMMA_DUAL=$d timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dual$d.log 2>&1
echo "== dual $d"; head -n 1 gpurun_out/prof_dual$d.log; grep -E "conv2d " gpurun_out/prof_dual$d.log | head -n 20
done
HOIG_UMMA_2CTA=2 timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_pair2.log 2>&1
echo "== dual 1 pair 2"; head -n 1 gpurun_out/prof_pair2.log; grep -E "conv2d " gpurun_out/prof_pair2.log | head -n 20
