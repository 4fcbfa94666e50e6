mkdir -p gpurun_out
cp hoig_b200/_C/libhoig_b200.so /tmp/lib_backup.so
for v in current dc76e90 current; do
cp hoig_b200/_C/variants/lib_$v.so hoig_b200/_C/libhoig_b200.so
timeout 600 python -c "
import ctypes, runpy, sys
sys.path.insert(0, '.')
import hoig_b200._lib as L
lib = ctypes.CDLL(L.LIB_PATH)
for k in list(L._SIGNATURES):
    if not hasattr(lib, k): L._SIGNATURES.pop(k)
sys.argv = ['profile_convs.py', '64', 'bf16']
runpy.run_path('scripts/profile_convs.py', run_name='__main__')
" > gpurun_out/prof_ab_$v.log 2>&1
echo "== variant $v"; head -n 1 gpurun_out/prof_ab_$v.log; grep -E "Cin512 Cout512|Cout1024|Cin512 Cout256|Cin128 Cout64 256|Cin256 Cout128 128" gpurun_out/prof_ab_$v.log | cut -c1-130
done
cp /tmp/lib_backup.so hoig_b200/_C/libhoig_b200.so
PT="python -m pytest -m gpu -q -x -p no:cacheprovider --timeout=300 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_umma.py > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"; tail -n 3 gpurun_out/t_umma.log | cut -c1-300
