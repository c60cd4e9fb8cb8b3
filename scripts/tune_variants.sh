#!/bin/bash
# tuning helper: GPU parity tests, then bench every libdiral_env*.so variant present (optionally with
# DIRAL_SMEM_PER_CTA settings from $PADS), then one ncu capture of the default build
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in diral_b200/libdiral_env*.so; do
  for pad in ${PADS:-0}; do
    DIRAL_SMEM_PER_CTA=$pad DIRAL_ENV_LIB=$PWD/$lib python bench.py --steps 300 --warmup 5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$lib pad=$pad', d['value'], d['ms_per_step'], d['roofline']['frac'], d['extra']['value_l2_resident_no_flush'], d['e2e']['value'])"
  done
done
ncu --set full --clock-control none --import-source on -k regex:step_group -s 105 -c 2 -o gpurun_out/prof_tune python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b2.log 2>&1
