#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_attention
timeout 300 python tools/one_op.py attn 8 8 4096 4096 40 | tail -1
timeout 300 python tools/one_op.py attn 8 8 1024 1024 80 | tail -1
timeout 300 python tools/one_op.py attn 8 8 256 256 160 | tail -1
timeout 300 python tools/one_op.py attn 8 8 4096 77 40 | tail -1
timeout 300 python tools/one_op.py attn 2 8 9216 9216 40 | tail -1
