#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_attention test_gpu_models
timeout 300 python tools/one_op.py attn 8 8 4096 4096 40 | tail -1
timeout 300 python tools/one_op.py attn 8 8 1024 1024 80 | tail -1
timeout 300 python tools/one_op.py attn 8 8 4096 77 40 | tail -1
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
timeout 600 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_all.log 2>&1
echo "bench rc=$? $(grep -o '"unet_step_ms": [0-9.]*' gpurun_out/bench_all.log) $(grep -o '"value": [0-9.]*' gpurun_out/bench_all.log | head -1)"
