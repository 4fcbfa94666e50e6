#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_umma.py tests/test_gpu_ops.py tests/test_gpu_generator.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3
for sh in s2_64 convT128 c128_64 stem heads; do
  for env in "HOIG_UMMA_FAST_EPI=0" "HOIG_UMMA_FAST_EPI=1"; do
    echo -n "$env  "; env $env python scripts/one_conv.py $sh 20 2>&1 | tail -n 1
  done
done
for m in 0 2 1 0 1; do
  HOIG_UMMA_FAST_EPI=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('fast_epi=$m', round(b['value'],1), round(b['e2e']['value'],1), round(b['roofline']['frac'],4), b['clocks']['sm_mhz'])"
done
} > gpurun_out/fast.log 2>&1
cat gpurun_out/fast.log
