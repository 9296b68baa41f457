#!/bin/bash
# Ablation record: the default one-MUFU sigmoid/tanh (tanh.approx.f32, ~2^-11) against ex2.approx + rcp.approx (~2^-22).
# Builds: python -m cesm_emulator_b200.build ; CESM_LIB_VARIANT=precise CESM_NVCC_EXTRA=-DCESM_PRECISE_SIGMOID python -m cesm_emulator_b200.build
for v in "" precise; do
  echo "=== build variant: ${v:-default (one-MUFU tanh.approx)} ==="
  for seed in 5 6 7; do
    CESM_LIB_VARIANT=$v SEED=$seed python tools/parity_report.py 2 3 128 128 2>/dev/null | sed -n 2,4p
  done
  CESM_LIB_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-comparator --no-extra --no-kernel-pass 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train step', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'samples/s')"
done
