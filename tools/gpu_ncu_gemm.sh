#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
run() { # name args...
  name=$1; shift
  python tools/one_gemm.py "$@" > gpurun_out/$name.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 2 -c 1 -f -o gpurun_out/$name python tools/one_gemm.py "$@" > gpurun_out/$name.ncu.log 2>&1
  echo "$name rc=$?"
}
run ncu_geglu_pair 32768 2560 320 0 1 1 0 1
run ncu_smallk_pair 32768 320 320 1 0 1
