#!/bin/bash
# Times the self-attention kernel built with each SDB_ATTN_VARIANT (measurement only; the variant libraries live in gpurun_out/).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PK=stable-diffusion-from-scratch_b200
for v in 0 1 2 3; do
  so=gpurun_out/libsdb200_v$v.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DSDB_ATTN_VARIANT=$v -shared -o $so $PK/csrc/*.cu -lcudart_static -ldl -lrt -lpthread || exit 1
  SDB200_LIB=$PWD/$so python tools/one_op.py attn 8 8 4096 4096 40 | tail -1 | sed "s/^/variant $v: /"
  rm -f $so
done
