PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT -s tests/test_gpu_halo.py 2>&1 | grep "attn_combine tc\|passed\|failed"
timeout 600 python scripts/profile_convs.py 64 f16 2>&1 | grep -E "forward|attn_combine"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:attn_combine --csv --log-file gpurun_out/r2_attn_launches.csv python scripts/profile_convs.py 8 f16 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_attn_launches.csv')) if len(r)>5]
h=rows[0]; i_val=h.index('Metric Value')
print([r[i_val] for r in rows[1:][-9:]])
PY
