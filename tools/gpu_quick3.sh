#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_bandwidth test_gpu_models
for a in "ln 32768 320" "ln 8192 640" "ln 2048 1280"; do
  timeout 300 python tools/one_op.py $a | tail -1
  SDB200_LN=regs timeout 300 python tools/one_op.py $a | tail -1 | sed 's/^/  (regs) /'
done
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
timeout 600 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_all.log 2>&1
echo "bench rc=$? $(grep -o '"unet_step_ms": [0-9.]*' gpurun_out/bench_all.log) $(grep -o '"value": [0-9.]*' gpurun_out/bench_all.log | head -1)"
