#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 600 python -m pytest tests/test_gpu_bandwidth.py -m gpu -q -x -p no:cacheprovider > gpurun_out/test_bw.log 2>&1; echo "bw tests rc=$? $(tail -1 gpurun_out/test_bw.log)"
cap() { # name kernel-regex count args...
  name=$1; shift; rx=$1; shift; cnt=$1; shift
  timeout 300 python tools/one_op.py "$@" > gpurun_out/$name.plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c $cnt -f -o gpurun_out/$name python tools/one_op.py "$@" > gpurun_out/$name.ncu.log 2>&1
  echo "$name rc=$? $(cat gpurun_out/$name.plain.log | tail -1)"
}
cap ncu_gnc_big gn_ 1 gn 8 4096 320 1
cap ncu_gnc_small gn_ 1 gn 8 64 1280 1
cap ncu_selfattn tc_attention 1 attn 8 8 4096 4096 40
cap ncu_ln layernorm 1 ln 32768 320
for cs in 16 8 4 2; do
  SDB200_GN_CS=$cs ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_ -s 4 -c 1 python tools/one_op.py gn 8 4096 320 1 2>&1 | grep -E "gpu__time|gn_cluster|gn_stats" | tr '\n' ' '; echo " cs=$cs big"
  SDB200_GN_CS=$cs ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_ -s 4 -c 1 python tools/one_op.py gn 8 64 1280 1 2>&1 | grep -E "gpu__time|gn_cluster|gn_stats" | tr '\n' ' '; echo " cs=$cs small"
done
