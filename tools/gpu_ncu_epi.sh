#!/bin/bash
# ncu source-level captures of two epilogue-paced contractions (GEGLU K=320, bf16-out K=320); measurement only
mkdir -p gpurun_out
python tools/one_op.py gemm 32768 2560 320 0 1 0 256 1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 6 -c 1 -f -o gpurun_out/full_epi_geglu python tools/one_op.py gemm 32768 2560 320 0 1 0 256 1 > gpurun_out/full_epi_geglu.log 2>&1; echo "rc=$?"
python tools/one_op.py gemm 32768 960 320 0 1 0 160 0 | tail -1
ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 6 -c 1 -f -o gpurun_out/full_epi_bf16 python tools/one_op.py gemm 32768 960 320 0 1 0 160 0 > gpurun_out/full_epi_bf16.log 2>&1; echo "rc=$?"
python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cat gpurun_out/configs.jsonl
