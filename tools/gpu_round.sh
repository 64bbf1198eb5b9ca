#!/bin/bash
# One GPU-box visit: tests, bench, launch list, full ncu captures of the top kernels.
mkdir -p gpurun_out
bash tools/gpu_tests.sh
python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_full.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref.log
python tools/profile_unet.py --batch 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_unet_b8.csv python tools/profile_unet.py --batch 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu launches rc=$?"
for k in tc_contract tc_attention_kernel gn_ layernorm_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 0 -c 3 -f \
      -o gpurun_out/full_$k python tools/profile_unet.py --batch 8 > gpurun_out/full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
