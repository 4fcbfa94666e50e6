mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_umma.py tests/test_gpu_generator.py > gpurun_out/t_umma.log 2>&1; echo "umma+gen rc=$?"; tail -n 14 gpurun_out/t_umma.log | cut -c1-300
for r in 1 2; do for m in 1 0; do
HOIG_UMMA_MMA_STATS=$m timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_ms${m}_$r.log 2>&1
echo "== mma_stats $m run $r"; head -n 1 gpurun_out/prof_ms${m}_$r.log; grep -E "Cin512 Cout512|convT k3 s2 Cin128|k7 s1 Cin64 Cout64|instnorm_apply   C64 256x256 N64 gb=0 res=0 ld=64->128|C512 32x32 N64 gb=0 res=1" gpurun_out/prof_ms${m}_$r.log | cut -c1-130
done; done
