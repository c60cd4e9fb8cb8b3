#!/bin/bash
# C3 slot time of the tuning builds of the lane-group kernel (resident warps per SM x warps per split environment)
for lib in diral_b200/libdiral_env.so diral_b200/libdiral_env_DIRAL_MIN_BLOCKS*.so; do
  echo "== $lib"
  DIRAL_ENV_LIB=$lib python scripts/bench_configs.py "C3 32x20" | cut -c1-200
  for ts in 0 1; do DIRAL_ENV_LIB=$lib DIRAL_TAIL_SPLIT=$ts python bench.py --steps 40 --warmup 5 --no-configs --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tail_split', $ts, 'ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'noflush', d['extra']['value_l2_resident_no_flush'])"; done
done
