#!/usr/bin/env python
"""Small driver for ncu: N launches of the predict kernel on a C<grid> x 72 matrix.
    python tools/profile_predict.py [--grid 90] [--iters 5] [--param name=value ...]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quickchem_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=90)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--rho", type=float, default=0.95)
ap.add_argument("--shuffle", action="store_true", help="random row order: worst-case lane coherence")
ap.add_argument("--model", default=None)
ap.add_argument("--param", action="append", default=[])
ap.add_argument("--sweep", default=None, help="name=v1,v2,...: time each setting")
a = ap.parse_args()
for kv in a.param:
    k, v = kv.split("=")
    capi.set_param(k, v)
x = synth.quick_features(synth.raw_fields(a.grid, rho=a.rho))
if a.shuffle:
    x = x[np.random.default_rng(0).permutation(x.shape[0])]
d = capi.DMatrix(x)  # before any booster exists: no pipelined prediction, only full-size launches below
b = capi.Booster(a.model or bench.booster_path())
out = capi.DeviceArray(x.shape[0])


def timeit(label):
    for _ in range(2):
        b.predict_device(d, out, exp10=True, scale=0.85)
    capi.synchronize()
    capi.timer_start()
    for _ in range(a.iters):
        b.predict_device(d, out, exp10=True, scale=0.85)
    ms = capi.timer_stop() / a.iters
    print(f"{label}: {ms:.3f} ms/launch  {x.shape[0] / ms / 1e3:.4g} Mcells/s", flush=True)
    return ms


if a.sweep:
    name, vals = a.sweep.split("=")
    for v in vals.split(","):
        capi.set_param(name, v)
        timeit(f"{name}={v}")
else:
    timeit("predict")
