#!/bin/bash
# tuning helper: C3 slot time (and, with TEST=1, the GPU parity suite) of every library variant named on the command line
for v in "$@"; do
  lib=diral_b200/libdiral_env$v.so
  echo "== $lib"
  DIRAL_ENV_LIB=$lib python scripts/bench_configs.py "${CFG:-C3 32x20}" | cut -c1-330
  if [ -n "$TEST" ]; then DIRAL_ENV_LIB=$lib python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
done
