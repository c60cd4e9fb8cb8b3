#!/bin/bash
# ncu capture of the one-CTA-per-env kernel at C4 (2048 envs x 128 UE x 64 resources)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:step_block -s 35 -c 1 -o gpurun_out/prof_block python scripts/bench_configs.py C4 > gpurun_out/b4.log 2>&1
tail -1 gpurun_out/b4.log
