#!/bin/bash
# One GPU-box visit: tests, bench, launch list, full ncu captures of the top kernels.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
[ -n "$SKIP_TESTS" ] || bash tools/gpu_tests.sh
python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_full.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref.log
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
timeout 600 python tools/bench_layers.py --what vae --batch 8 --variants 1 --json gpurun_out/layers_vae_b8.json > gpurun_out/layers_vae_b8.log 2>&1; echo "layers vae rc=$?"
python tools/profile_unet.py --batch 8 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_unet_b8.csv python tools/profile_unet.py --batch 8 > gpurun_out/prof_ncu.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:tc_contract --csv \
    --log-file gpurun_out/traffic_tc_unet_b8.csv python tools/profile_unet.py --batch 8 > gpurun_out/prof_traffic.log 2>&1
echo "ncu traffic rc=$?"
for k in tc_contract tc_attention_kernel gn_ layernorm_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 0 -c 3 -f \
      -o gpurun_out/full_$k python tools/profile_unet.py --batch 8 > gpurun_out/full_$k.log 2>&1
  echo "ncu full $k rc=$?"
done
