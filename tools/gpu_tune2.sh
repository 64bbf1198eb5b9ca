#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu_models.log 2>&1; echo "models rc=$? $(tail -1 gpurun_out/test_gpu_models.log)"
SDB200_TC_PLANS=0 timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_noplans.json > gpurun_out/layers_noplans.log 2>&1; echo "layers(noplans) rc=$?"
grep -E "variant|by entry|rel-L2" gpurun_out/layers_noplans.log
SDB200_TC_PLANS=0 timeout 1200 python tools/tune_tc.py --batch 8 --out gpurun_out/tune_tc_unet_b8.jsonl > gpurun_out/tune_tc.log 2>&1; echo "tune rc=$?"
tail -2 gpurun_out/tune_tc.log
