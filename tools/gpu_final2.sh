#!/bin/bash
# Round-2 end rehearsal: what the driver runs (single-process pytest -m gpu, smoke, both bench arms), the other BASELINE configs,
# per-layer tables, the library bar, then the profiling visit (tools/gpu_round2.sh: launch list, DRAM traffic, labelled ncu captures).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || echo "build failed"
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$? :: $(tail -1 gpurun_out/pytest_gpu_all.log)"
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? :: $(tail -1 gpurun_out/smoke.log | cut -c1-300)"
python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-1500
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
for c in c3 c5; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_$c.log 2>&1; echo "bench $c rc=$?"; tail -1 gpurun_out/bench_$c.log | cut -c1-600
done
timeout 600 python tools/bench_layers.py --batch 8 --variants 1 --json gpurun_out/layers_unet_b8.json > gpurun_out/layers_unet_b8.log 2>&1; echo "layers rc=$?"
timeout 600 python tools/bench_layers.py --what vae --batch 8 --variants 1 --json gpurun_out/layers_vae_b8.json > gpurun_out/layers_vae_b8.log 2>&1; echo "layers vae rc=$?"
timeout 900 python tests/library_bar.py > gpurun_out/r02_library_bar.txt 2> gpurun_out/library_bar.log; echo "library bar rc=$?"
for i in 1 2; do python tools/ab_unet_step.py default; SDB200_SKIP_SIDE_STREAM=0 SDB200_EMB_SIDE_STREAM=0 python tools/ab_unet_step.py no_side_streams; SDB200_XATTN=0 python tools/ab_unet_step.py key_walking_xattn; SDB200_CONV_IN_TC=0 python tools/ab_unet_step.py conv_in_simt; done 2>&1 | grep "UNet step" | tee gpurun_out/ab_unet_step_final.txt
[ -n "$SKIP_PROFILES" ] || bash tools/gpu_round2.sh
