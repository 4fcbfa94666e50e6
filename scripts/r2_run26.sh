PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT -s tests/test_gpu_halo.py tests/test_gpu_umma.py 2>&1 | grep "attn_combine tc\|passed\|failed\|FAILED"
timeout 900 $PT tests/test_gpu_generator.py 2>&1 | tail -2
for m in 1 0 1 0; do HOIG_ATTN_PHASE1_TC=$m timeout 900 python bench.py --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('phase1_tc=$m value',round(d['value'],1),'eager',round(d['roofline']['eager_ms_per_step'],2),'attn ms',round(d['kernel_time_share']['attn_combine']*d['roofline']['eager_ms_per_step'],3))"; done
