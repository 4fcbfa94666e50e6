for s in stem c128_64 convT128 s2_64 heads; do timeout 300 python scripts/one_conv.py $s 10; done
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py 2>&1 | tail -3
timeout 600 python scripts/profile_convs.py 64 f16 2>&1 | head -12 | cut -c1-150
