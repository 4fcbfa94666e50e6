mkdir -p gpurun_out
timeout 300 python scripts/halo_probe.py > gpurun_out/halo_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/halo_probe.log | cut -c1-400
