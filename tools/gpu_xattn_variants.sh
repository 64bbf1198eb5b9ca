#!/bin/bash
# which stage paces tc_attention_kv1_kernel: measurement builds variants/libsdb200_xa<N>.so (-DSDB_XATTN_DEBUG=N: 1 no output stores,
# 2 no MUFU, 3 Q tiles loaded once, 4 no P stores) against the shipped library, device time by CUDA-graph replay
mkdir -p gpurun_out
for shape in "16 8 4096 77 40" "8 8 4096 77 40"; do
  echo "base $(ONE_OP_GRAPH=1 python tools/one_op.py attn $shape 2>&1 | tail -3 | head -1)"
  for so in variants/libsdb200_xa*.so; do
    echo "$(basename $so) $(ONE_OP_GRAPH=1 SDB200_LIB=$PWD/$so python tools/one_op.py attn $shape 2>&1 | tail -3 | head -1)"
  done
done 2>&1 | tee gpurun_out/xattn_variants.txt
