#!/bin/bash
# GPU visit: tests, per-layer timings, then A/B of the step with each new mechanism switched off.
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
b() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  echo "bench $name rc=$? $(grep -o '"unet_step_ms": [0-9.]*' gpurun_out/bench_$name.log) $(grep -o '"value": [0-9.]*' gpurun_out/bench_$name.log | head -1)"
}
b all X=1
b nopdl SDB200_PDL=0
b gnsplit SDB200_GN=split
b noplans SDB200_TC_PLANS=0
for a in "gn 8 4096 320 1" "gn 8 1024 640 1" "gn 8 256 1280 1" "gn 8 64 1280 1" "gemm 32768 320 320 1 0 2 160 0" "gemm 32768 320 320 1 0 1 160 0" "conv 8 64 64 320 320 3 1 1 2 160" "conv 8 64 64 320 320 3 1 1 1 160"; do
  timeout 300 python tools/one_op.py $a 2>&1 | tail -1
done
