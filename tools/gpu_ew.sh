#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_tc_gemm test_gpu_tc_conv test_gpu_tc_large
for ew in 8 16 0; do
  SDB200_TC_EW=$ew timeout 600 python tools/bench_layers.py --batch 8 --variants 1 > gpurun_out/layers_ew$ew.log 2>&1; echo "ew=$ew rc=$?"
  head -1 gpurun_out/layers_ew$ew.log; grep "by entry point" gpurun_out/layers_ew$ew.log
done
for s in "32768 2560 320 0 1 0 256 1" "32768 960 320 0 1 0 160 0" "32768 320 320 1 0 0 160 0" "8192 5120 640 0 1 0 256 1" "2048 1280 1280 1 0 0 0 0"; do
  for ew in 8 16; do SDB200_TC_EW=$ew python tools/one_op.py gemm $s | tail -1 | sed "s/^/ew=$ew /"; done
done
