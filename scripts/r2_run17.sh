PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_ops.py tests/test_gpu_generator.py 2>&1 | tail -3
for b in 1 0 1 0; do HOIG_BATCH2=$b timeout 900 python bench.py --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('batch2=$b value',round(d['value'],1),'eager',round(d['roofline']['eager_ms_per_step'],2),'conv',round(d['roofline']['conv_ms_per_step'],2), {k:round(v*d['roofline']['eager_ms_per_step'],2) for k,v in d['kernel_time_share'].items()})"; done
