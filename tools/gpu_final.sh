#!/bin/bash
# Round-end rehearsal: what the driver runs (single-process pytest -m gpu, smoke, both bench arms), then the profile round.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? :: $(tail -1 gpurun_out/pytest_gpu_all.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? :: $(tail -1 gpurun_out/smoke.log | cut -c1-300)"
SKIP_TESTS=1 bash tools/gpu_round.sh
