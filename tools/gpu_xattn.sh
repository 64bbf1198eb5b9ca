#!/bin/bash
# short-key cross-attention kernel (tc_attention_kv1_kernel): parity, then old/new device time (CUDA-graph replay of 20 calls)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q > gpurun_out/test_gpu_attention.log 2>&1; echo "test_gpu_attention rc=$? :: $(tail -1 gpurun_out/test_gpu_attention.log)"
for shape in "8 8 4096 77 40" "8 8 1024 77 80" "8 8 256 77 160" "16 8 4096 77 40" "16 8 1024 77 80" "1 8 4096 77 40"; do
  for x in 0 1; do
    echo "XATTN=$x $(ONE_OP_GRAPH=1 SDB200_XATTN=$x timeout 120 python tools/one_op.py attn $shape 2>&1 | tail -3 | tr '\n' ' ')"
  done
done | tee gpurun_out/xattn_times.txt
