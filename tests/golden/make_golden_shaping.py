#!/usr/bin/env python
"""Pins oracle/shaping.py::ia_penalty_sum to the reference's utils/misc.py::calculate_ia_penalty (run in
the build container: needs /root/reference).  Writes tests/golden/shaping_ia_sums.json."""
import json
import os
import sys

import numpy as np

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from utils.misc import calculate_ia_penalty  # noqa: E402  (the reference function)

rs = np.random.RandomState(9)
cases = []
for _ in range(64):
    ia = (rs.randint(0, 40, size=100) * (rs.rand(100) < 0.3)).astype(int).tolist()
    cases.append({"ia": ia, "sum": int(calculate_ia_penalty(ia))})
cases.append({"ia": [0] * 100, "sum": int(calculate_ia_penalty([0] * 100))})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shaping_ia_sums.json")
json.dump(cases, open(out, "w"))
print(out, len(cases))
