#!/usr/bin/env python
"""Kernel-shape experiments of the two-level kernel (experiment build: make -C quickchem_b200/csrc exp;
QCOH_LIB=quickchem_b200/libqcoh_exp.so): trees in flight x resident CTAs per SM x texture-pipe mask, per booster.
    QCOH_LIB=quickchem_b200/libqcoh_exp.so python tools/sweep_shapes.py --models 10x6,10x10 --grid 180"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickchem_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=180)
ap.add_argument("--models", default="10x6,10x10")
ap.add_argument("--shapes", default="0:0:0,6:5:0x100,6:5:0x14,4:6:0xA,4:6:0x100,3:6:0x100,3:6:0x2,2:6:0x100,2:6:0x2,5:6:0x100,5:6:0xA,"
                                    "4:7:0x100,4:7:0xA,3:7:0x100,3:7:0x2,2:7:0x100,8:4:0xEE")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
x = synth.quick_features(synth.raw_fields(a.grid))
d = capi.DMatrix(x)
out = capi.DeviceArray(x.shape[0])
peak = 6525.2
for m in a.models.split(","):
    b = capi.Booster(os.path.join(ROOT, "build", f"oh_booster_{m}.model"))
    for shp in a.shapes.split(","):
        ilp, minb, mask = (int(v, 0) for v in shp.split(":"))
        capi.set_param("duo", 1 if ilp else -1)
        capi.set_param("ilp", ilp); capi.set_param("minb", minb); capi.set_param("duo_mask", mask)
        for persist in ((0, 1) if ilp == 0 else (0,)):
            capi.set_param("persist", persist)
            for _ in range(3):
                b.predict_device(d, out, exp10=True, scale=0.85)
            capi.synchronize(); capi.timer_start()
            for _ in range(a.iters):
                b.predict_device(d, out, exp10=True, scale=0.85)
            ms = capi.timer_stop() / a.iters
            gbs = x.shape[0] * 112 / 1e9 / (ms * 1e-3)
            print(json.dumps(dict(model=m, ilp=ilp, minb=minb, mask=hex(mask), persist=persist, kernel=capi.last_predict_kernel(),
                                  ms=round(ms, 4), cells_per_s=x.shape[0] / ms * 1e3, hbm_frac=round(gbs / peak, 4))), flush=True)
    b.free()
