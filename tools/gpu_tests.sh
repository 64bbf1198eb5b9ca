#!/bin/bash
# Runs every -m gpu test file in its own process (a faulting kernel poisons only its own CUDA context).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
for f in ${@:-test_gpu_bandwidth test_gpu_simt test_gpu_tc_gemm test_gpu_tc_conv test_gpu_tc_large test_gpu_tc_epilogue test_gpu_attention test_gpu_clip test_gpu_models test_gpu_full_size}; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s -p no:cacheprovider --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$? :: $(tail -1 gpurun_out/$f.log)"
done
