#!/bin/bash
# Re-measure the launch-plan table of the UNet contractions, then A/B the UNet step with the old and the new table.
mkdir -p gpurun_out
step() { python bench.py --steps 1 --warmup 3 --skip-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],3), 'img/s  unet_step_ms', round(d['unet_step_ms'],3))"; }
step old-plans
cp stable-diffusion-from-scratch_b200/tc_plans.py gpurun_out/tc_plans_old.py
SECONDS=0
timeout 500 python tools/tune_tc.py --batch 8 --out gpurun_out/tune_tc_unet_b8.jsonl > gpurun_out/tune_tc.log 2>&1; echo "sweep rc=$? in ${SECONDS}s"
python tools/make_tc_plans.py gpurun_out/tune_tc_unet_b8.jsonl tools/sweeps/tune_tc_vae_b8.jsonl
cp stable-diffusion-from-scratch_b200/tc_plans.py gpurun_out/tc_plans_new.py
step new-plans
cp gpurun_out/tc_plans_old.py stable-diffusion-from-scratch_b200/tc_plans.py
step old-plans
cp gpurun_out/tc_plans_new.py stable-diffusion-from-scratch_b200/tc_plans.py
step new-plans
