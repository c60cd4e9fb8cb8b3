#!/bin/bash
python -m pytest tests -m gpu -x -q -k "block or c4_ or c5_" 2>&1 | tail -2
for lib in diral_b200/libdiral_env*.so; do
  echo "== $lib"
  DIRAL_ENV_LIB=$PWD/$lib python scripts/bench_configs.py C4 2>&1 | python -c "import json,sys; [print(d['config'], round(d['us_per_slot'],1), round(d['roofline_frac'],3)) for d in map(json.loads, sys.stdin)]"
  DIRAL_ENV_LIB=$PWD/$lib python scripts/bench_configs.py "C5 64" 2>&1 | python -c "import json,sys; [print(d['config'], round(d['us_per_slot'],1), round(d['roofline_frac'],3)) for d in map(json.loads, sys.stdin)]"
  DIRAL_ENV_LIB=$PWD/$lib python scripts/bench_configs.py "C5 256" 2>&1 | python -c "import json,sys; [print(d['config'], round(d['us_per_slot'],1), round(d['roofline_frac'],3)) for d in map(json.loads, sys.stdin)]"
done
