#!/bin/bash
# shared-memory-resident halo for 3x3 convs: parity, then halo on / off and one-CTA kernel at the UNet / VAE shapes (device time, graph replay)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -x -q > gpurun_out/test_gpu_tc_conv.log 2>&1; echo "test_gpu_tc_conv rc=$? :: $(tail -1 gpurun_out/test_gpu_tc_conv.log)"
tail -30 gpurun_out/test_gpu_tc_conv.log | grep -E "Error|assert|FAILED" | head -10
#            NB H W Cin Cout ks stride res variant bn
for shape in "8 64 64 320 320 3 1 1" "8 64 64 640 320 3 1 0" "8 32 32 640 640 3 1 1" "8 32 32 1280 640 3 1 0" "8 16 16 1280 1280 3 1 1" "8 16 16 2560 1280 3 1 0" "2 512 512 128 128 3 1 1" "2 256 256 256 256 3 1 1" "2 128 128 512 512 3 1 1"; do
  for cfg in "1 2 160" "0 2 160" "0 1 160" "1 2 256" "0 2 256" "1 2 128"; do
    set -- $cfg
    echo "halo=$1 variant=$2 bn=$3 :: conv $shape :: $(ONE_OP_GRAPH=1 SDB200_TC_HALO=$1 timeout 120 python tools/one_op.py conv $shape $2 $3 2>&1 | grep -E "graph replay|Error|error" | head -1)"
  done
done 2>&1 | tee gpurun_out/halo_times.txt
