#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
rm -f gpurun_out/libsdb200_trace.so
python tools/trace_pair.py 32768 320 320 1 0 160 > gpurun_out/trace_smallk.txt 2>&1; echo "trace1 rc=$?"
python tools/trace_pair.py 32768 960 320 0 1 160 > gpurun_out/trace_qkv.txt 2>&1; echo "trace2 rc=$?"
python tools/trace_pair.py 32768 2560 320 0 1 256 1 > gpurun_out/trace_geglu.txt 2>&1; echo "trace3 rc=$?"
python tools/trace_pair.py 32768 320 2880 1 0 160 > gpurun_out/trace_k2880.txt 2>&1; echo "trace4 rc=$?"
python tools/trace_pair.py 32768 1280 5760 0 0 256 > gpurun_out/trace_big.txt 2>&1; echo "trace5 rc=$?"
rm -f gpurun_out/libsdb200_trace.so
SDB200_TC_PLANS=0 timeout 1200 python tools/tune_tc.py --batch 8 --out gpurun_out/tune_tc_unet_b8.jsonl > gpurun_out/tune_tc.log 2>&1; echo "tune rc=$?"
tail -2 gpurun_out/tune_tc.log
