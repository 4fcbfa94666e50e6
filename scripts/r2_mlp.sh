#!/bin/bash
mkdir -p gpurun_out
{
for sh in mlp128 mlp896; do
  for env in "" "HOIG_UMMA_DEBUG=1" "HOIG_UMMA_DEBUG=8" "HOIG_UMMA_2CTA=0" "HOIG_UMMA_DUAL=0" "HOIG_UMMA_BRES=0" "HOIG_UMMA_PREFETCH=0"; do
    echo -n "$env  "; env $env python scripts/one_conv.py $sh 20 2>&1 | tail -n 1
  done
done
} > gpurun_out/mlp.log 2>&1
cat gpurun_out/mlp.log
