#!/bin/bash
# round-2 timelines of the pair kernel with the TMA epilogue at the bandwidth- and latency-bound shapes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
rm -f gpurun_out/libsdb200_trace.so
python tools/trace_pair.py 32768 320 320 1 0 160 > gpurun_out/trace2_smallk.txt 2>&1; echo "trace1 rc=$?"
python tools/trace_pair.py 2048 1280 1280 1 0 160 > gpurun_out/trace2_m2048.txt 2>&1; echo "trace2 rc=$?"
python tools/trace_pair.py 32768 320 1280 1 0 160 > gpurun_out/trace2_ffout.txt 2>&1; echo "trace3 rc=$?"
python tools/trace_pair.py 8192 640 640 1 0 160 > gpurun_out/trace2_m8192.txt 2>&1; echo "trace4 rc=$?"
rm -f gpurun_out/libsdb200_trace.so
