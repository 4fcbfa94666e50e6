for m in 1 2; do echo "== 2CTA=$m"; for s in convT128 s2_64 c128_64; do HOIG_UMMA_2CTA=$m timeout 300 python scripts/one_conv.py $s 10; done; done
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench_mid.log 2>&1; tail -1 gpurun_out/r2_bench_mid.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'burst',d['roofline']['frac_of_burst_peak'],'conv_ms',d['roofline']['conv_ms_per_step'],'eager',d['roofline']['eager_ms_per_step'])
print(d['kernel_time_share'])"
