#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity suite (both kernels, all three modes)
SEL='kat3_design6_ch_d3 or n48x10_my_step-block-fused or c3_32x20_step_design-group-fused or toy4x3_shipped_T80-group-fused or n70x16_ch_d3-block-split'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|hazard|Race" | head -12
done
