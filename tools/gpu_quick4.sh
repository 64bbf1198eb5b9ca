#!/bin/bash
mkdir -p gpurun_out
./variants/xu_rate > gpurun_out/xu_rate.txt 2>&1; cat gpurun_out/xu_rate.txt
for s in "32768 2560 320 0 1 1 256 1" "8192 5120 640 0 1 1 256 1" "2048 10240 1280 0 1 1 256 1"; do python tools/one_op.py gemm $s | tail -1; done 2>&1 | tee gpurun_out/geglu_new.txt
bash tools/gpu_tests.sh test_gpu_tc_gemm test_gpu_models
