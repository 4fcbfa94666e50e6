# Refresh of the conv evidence after the last kernel changes: launch list of the bench command, conv DRAM traffic, --set full conv tables.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma_kernel|conv_halo_kernel" -c 400 --csv --log-file gpurun_out/r02_conv_traffic.csv env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
full() {
  name=$1; regex=$2; shift 2
  timeout 1200 ncu --set full --clock-control none -k regex:"$regex" "$@" > gpurun_out/r02_ncu_$name.log 2>&1; rc=$?
  rep=/tmp/r02_$name.ncu-rep
  if [ -f $rep ]; then ncu -i $rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null; rm -f $rep; fi
  echo "ncu $name rc=$rc"
}
full conv_umma_a "conv_umma_kernel" -c 44 -o /tmp/r02_conv_umma_a -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
full conv_umma_b "conv_umma_kernel" --launch-skip 88 -c 44 -o /tmp/r02_conv_umma_b -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
gzip -f gpurun_out/r02_launches_bench.csv gpurun_out/r02_conv_traffic.csv
du -sh gpurun_out
