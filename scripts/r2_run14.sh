PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT -s tests/test_gpu_halo.py 2>&1 | grep -v "^E   .*tensor\|^E    *\[" | tail -25
timeout 900 $PT tests/test_gpu_generator.py 2>&1 | tail -3
timeout 600 python scripts/profile_convs.py 64 f16 2>&1 | grep -E "forward|attn_combine|halo|replicate"
HOIG_ATTN_TC=0 timeout 600 python scripts/profile_convs.py 64 f16 2>&1 | grep -E "forward|attn_combine"
