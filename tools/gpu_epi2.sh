#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_tc_gemm test_gpu_tc_conv test_gpu_tc_large test_gpu_models
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 > gpurun_out/layers_epi2.log 2>&1; echo "layers rc=$?"
head -1 gpurun_out/layers_epi2.log; grep "by entry point" gpurun_out/layers_epi2.log
for s in "32768 2560 320 0 1 0 256 1" "32768 960 320 0 1 0 160 0" "32768 320 320 1 0 0 160 0" "8192 5120 640 0 1 0 256 1" "2048 1280 1280 1 0 0 0 0"; do
  python tools/one_op.py gemm $s | tail -1
done
python bench.py --steps 1 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], 'img/s  unet_step_ms', d['unet_step_ms'], 'roofline', d['roofline']['frac'])"
