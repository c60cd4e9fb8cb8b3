#!/bin/bash
# compute-sanitizer passes over a slice of the GPU parity suite that reaches every slot kernel and every special path:
# lane-group kernel (all three modes, fused rollout, split environments, 32-bit fallback slabs after 2 047 slots are out
# of reach of a sanitizer run -- the stale tests below use the sparse-highway configs that fall back within ~1 100 slots),
# pair kernel (33..64 vehicles), round-1 block kernel (shared-memory and scratch keys), row kernel (flag-ordered merges,
# wide keys), compact host format (chunked, zero-copy, streamed records with TMA bulk stores and device-raised flags,
# diral_step_host_begin / _wait).
SEL='kat3_design6_ch_d3 or n48x10_my_step or c3_32x20_step_design-group-fused or toy4x3_shipped_T80-group-fused or n70x16_ch_d3 or n130x40_my_step-row-fused or n160x70 or n200x90_ch-block_v1 or c5_100x50-row or state_vpd1 or state_no_piggy_direct or state_real_action_vpd1'
SEL2='fused_rollout_equals_slot_by_slot and 16-8 or split_environment_kernel_reproduces_fixtures and c3_32x20_ch_d3 or compact_host_format_is_bit_identical and 13-7 or compact_host_format_is_bit_identical and 32-20-2048 or begin_wait'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Race" | head -12
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -k "$SEL2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Race" | head -12
done
echo "== memcheck, stale entries (32-bit key fallback: pair / row / block_v1 kernels)"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "stale_entries_take and 140" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error" | head -6
