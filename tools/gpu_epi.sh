#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh
rm -f gpurun_out/libsdb200_trace.so
python tools/trace_pair.py 32768 320 320 1 0 160 > gpurun_out/trace_smallk.txt 2>&1; echo "trace1 rc=$?"
python tools/trace_pair.py 32768 2560 320 0 1 256 1 > gpurun_out/trace_geglu.txt 2>&1; echo "trace3 rc=$?"
rm -f gpurun_out/libsdb200_trace.so
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
b() { # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  echo "bench $name rc=$? $(grep -o '"unet_step_ms": [0-9.]*' gpurun_out/bench_$name.log) $(grep -o '"value": [0-9.]*' gpurun_out/bench_$name.log | head -1)"
}
b all X=1
b noplans SDB200_TC_PLANS=0
