#!/usr/bin/env python
"""Time XGDMatrixCreateFromMat / XGBoosterPredict / XGDMatrixFree separately (pinned host X)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from quickchem_b200 import capi, synth
ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=180)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
b = capi.Booster(bench.booster_path())
x = synth.quick_features(synth.raw_fields(a.grid))
hx = capi.pinned_empty(x.shape); hx[:] = x
print("rows", x.shape[0], "GB", x.nbytes / 1e9, flush=True)
import os as _os
cases = ((0, 0, hx, "pinned"), (1, 0, hx, "pinned"), (1, 1 << 19, hx, "pinned"), (0, 0, x, "pageable"), (1, 0, x, "pageable"))
if _os.environ.get("PAGEABLE_ONLY"):
    cases = ((1, 0, hx, "pinned"), (1, 0, x, "pageable"), (1, 1 << 18, x, "pageable"), (1, 1 << 17, x, "pageable"))
for spec, chunk, src, name in cases:
    capi.set_param("speculate", spec); capi.set_param("chunk_rows", chunk)
    print(name, flush=True)
    for it in range(a.iters):
        t0 = time.perf_counter(); d = capi.DMatrix(src); t1 = time.perf_counter()
        n, p = b.predict_raw(d); t2 = time.perf_counter()
        d.free(); t3 = time.perf_counter()
        print(f"spec={spec} chunk={chunk} it={it}: create {1e3*(t1-t0):7.2f}  predict {1e3*(t2-t1):7.2f}  free {1e3*(t3-t2):6.2f}  total {1e3*(t3-t0):7.2f} ms", flush=True)
