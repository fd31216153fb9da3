#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz from the CPU oracle:  python tests/golden/make_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import cases  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.build()
orc.set_mode(0)
out = {}
for name in cases.CASES:
    for k, v in cases.oracle_outputs(orc, name).items():
        out[f"{name}/{k}"] = v
    print(name, {k.split("/")[1]: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items() if k.startswith(name + "/")})
np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")
