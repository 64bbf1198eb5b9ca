#!/bin/bash
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
b nocolstats SDB200_COLSTATS=0
b nopdl SDB200_PDL=0
