#!/bin/bash
# GPU visit: full GPU test suite, per-layer timings, launch-plan sweep, source-level ncu captures of the slow classes.
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_unet_b8.log
timeout 900 python tools/tune_tc.py --batch 8 --out gpurun_out/tune_tc_unet_b8.jsonl > gpurun_out/tune_tc.log 2>&1; echo "tune rc=$?"
tail -3 gpurun_out/tune_tc.log
cap() { # name kernel-regex args...
  name=$1; shift; rx=$1; shift
  timeout 300 python tools/one_op.py "$@" > gpurun_out/$name.plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o gpurun_out/$name python tools/one_op.py "$@" > gpurun_out/$name.ncu.log 2>&1
  echo "$name rc=$? $(cat gpurun_out/$name.plain.log | tail -1)"
}
cap ncu_smallk_pair tc_contract gemm 32768 320 320 1 0 2 160 0
cap ncu_qkv_pair tc_contract gemm 32768 960 320 0 1 2 160 0
cap ncu_conv320_pair tc_contract conv 8 64 64 320 320 3 1 1 2 160
cap ncu_xattn tc_attention attn 8 8 4096 77 40
cap ncu_gn gn_ gn 8 4096 320 1
python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_quick.log
