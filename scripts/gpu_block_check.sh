#!/bin/bash
# parity of the whole GPU suite, the block-kernel configs, and one ncu --set full capture at C4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in C4 "C5 64" "C5 256"; do timeout 120 python scripts/bench_configs.py "$c" 2>&1 | cut -c1-230; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_block -s 35 -c 1 -o gpurun_out/${1:-prof_block} python scripts/bench_configs.py C4 > gpurun_out/b4.log 2>&1
tail -1 gpurun_out/b4.log | cut -c1-200
