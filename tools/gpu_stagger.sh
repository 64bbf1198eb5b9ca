#!/bin/bash
# stagger of the two co-resident CTAs of the one-CTA kernel: device time (graph replay) at the UNet's big one-CTA shapes, then the UNet step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
for ns in 0 2000 4000 6000 9000; do
  echo "stagger=$ns conv320@64 $(ONE_OP_GRAPH=1 SDB200_TC_STAGGER_NS=$ns python tools/one_op.py conv 8 64 64 320 320 3 1 1 1 160 2>&1 | grep graph | cut -c14-) | conv640->320@64 $(ONE_OP_GRAPH=1 SDB200_TC_STAGGER_NS=$ns python tools/one_op.py conv 8 64 64 640 320 3 1 0 1 160 2>&1 | grep graph | cut -c14-) | conv640@32 $(ONE_OP_GRAPH=1 SDB200_TC_STAGGER_NS=$ns python tools/one_op.py conv 8 32 32 640 640 3 1 1 1 160 2>&1 | grep graph | cut -c14-) | gemm 32768x320x1280 $(ONE_OP_GRAPH=1 SDB200_TC_STAGGER_NS=$ns python tools/one_op.py gemm 32768 320 1280 1 0 1 160 0 2>&1 | grep graph | cut -c14-) | gemm 8192x640x2560 $(ONE_OP_GRAPH=1 SDB200_TC_STAGGER_NS=$ns python tools/one_op.py gemm 8192 640 2560 1 0 1 160 0 2>&1 | grep graph | cut -c14-)"
done | tee gpurun_out/stagger_times.txt
for ns in 0 4000 6000 0 4000 6000; do SDB200_TC_STAGGER_NS=$ns python tools/ab_unet_step.py stagger_$ns; done 2>&1 | grep "UNet step" | tee -a gpurun_out/stagger_times.txt
