PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py tests/test_gpu_geometry_fixture.py 2>&1 | tail -4
timeout 600 python scripts/profile_convs.py 64 f16 2>&1 | head -8 | cut -c1-150
timeout 900 python bench.py --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv',round(d['roofline']['conv_ms_per_step'],2),'eager',round(d['roofline']['eager_ms_per_step'],2),d['clocks']['sm_mhz'])"
