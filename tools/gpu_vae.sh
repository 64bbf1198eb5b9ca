#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_models
timeout 600 python tools/bench_layers.py --what vae --batch 8 --variants 1 --json gpurun_out/layers_vae_b8.json > gpurun_out/layers_vae_b8.log 2>&1; echo "layers vae rc=$?"
grep -E "variant|by entry" gpurun_out/layers_vae_b8.log
SDB200_TC_PLANS=0 timeout 1200 python tools/tune_tc.py --what vae --batch 8 --iters 6 --out gpurun_out/tune_tc_vae_b8.jsonl > gpurun_out/tune_tc_vae.log 2>&1; echo "tune vae rc=$?"
tail -3 gpurun_out/tune_tc_vae.log
