#!/bin/bash
# Times prebuilt attention-kernel variants (variants/libsdb200_v<VARIANT>_p<POLY>.so, built on the dev box) on the UNet's self-attention shapes.
mkdir -p gpurun_out
for so in variants/libsdb200_*.so; do
  for shape in "8 8 4096 4096 40" "8 8 1024 1024 80" "8 8 4096 77 40"; do
    SDB200_LIB=$PWD/$so python tools/one_op.py attn $shape | tail -2 | tr '\n' ' ' | sed "s|^|$(basename $so) |"; echo
  done
done 2>&1 | tee gpurun_out/attn_variants.log
