#!/bin/bash
# A/B of the shipped library against prebuilt variants (variants/libsdb200_*.so) on the UNet's self-attention shapes, device time by graph replay
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q > gpurun_out/test_gpu_attention.log 2>&1; echo "test_gpu_attention rc=$? :: $(tail -1 gpurun_out/test_gpu_attention.log)"
for shape in "8 8 4096 4096 40" "8 8 1024 1024 80" "8 8 256 256 160"; do
  echo "shipped $(ONE_OP_GRAPH=1 python tools/one_op.py attn $shape 2>&1 | tail -3 | tr '\n' ' ')"
  for so in variants/libsdb200_*.so; do
    echo "$(basename $so) $(ONE_OP_GRAPH=1 SDB200_LIB=$PWD/$so python tools/one_op.py attn $shape 2>&1 | tail -3 | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/attn_ab.txt
