#!/bin/bash
# quick GPU visit: contraction tests + micro-benchmarks + per-layer timings of both kernel variants
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
for f in test_gpu_tc_gemm test_gpu_tc_conv test_gpu_tc_large; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f rc=$? :: $(tail -1 gpurun_out/$f.log)"
done
timeout 600 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "bench_gemm rc=$?"
timeout 600 python tools/bench_layers.py --batch 8 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
