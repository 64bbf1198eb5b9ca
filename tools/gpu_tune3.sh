#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_tc_conv test_gpu_models
SDB200_TC_PLANS=0 timeout 1200 python tools/tune_tc.py --batch 8 --out gpurun_out/tune_tc_unet_b8.jsonl > gpurun_out/tune_tc.log 2>&1; echo "tune unet rc=$?"
SDB200_TC_PLANS=0 timeout 1200 python tools/tune_tc.py --what vae --batch 8 --iters 6 --out gpurun_out/tune_tc_vae_b8.jsonl > gpurun_out/tune_tc_vae.log 2>&1; echo "tune vae rc=$?"
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
timeout 600 python tools/bench_layers.py --what vae --batch 8 --variants 1 --json gpurun_out/layers_vae_b8.json > gpurun_out/layers_vae_b8.log 2>&1
grep -E "variant|by entry" gpurun_out/layers_vae_b8.log
