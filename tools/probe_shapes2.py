#!/usr/bin/env python
"""Launch shape x trees-per-launch probe for forests with many and / or shallow trees (experiment build).
    QCOH_LIB=quickchem_b200/libqcoh_exp.so python tools/probe_shapes2.py"""
import json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickchem_b200 import capi, synth, xgbmodel  # noqa: E402

x = synth.quick_features(synth.raw_fields(180))
d = capi.DMatrix(x)
out = capi.DeviceArray(x.shape[0])
tmp = tempfile.mkdtemp()


def t(b, iters=4):
    for _ in range(2):
        b.predict_device(d, out)
    capi.synchronize(); capi.timer_start()
    for _ in range(iters):
        b.predict_device(d, out)
    return capi.timer_stop() / iters


def model(name):
    trees, depth = (int(v) for v in name.split("x"))
    p = os.path.join(ROOT, "build", f"oh_booster_{name}.model")
    if os.path.exists(p):
        return p
    f = synth.replicate_forest(xgbmodel.read_legacy_binary(os.path.join(ROOT, "build", f"oh_booster_100x{depth}.model")), trees // 100, seed=trees + depth)
    p = os.path.join(tmp, f"{name}.model")
    xgbmodel.write_legacy_binary(f, p)
    return p


shapes = ("0:0:0", "4:6:0x100", "4:6:0xA", "5:6:0xA", "5:6:0x100", "6:5:0x100")
for m in sys.argv[1:] or ("100x6", "100x10", "100x18", "500x6", "500x10", "500x14"):
    b = capi.Booster(model(m))
    for shp in shapes:
        ilp, minb, mask = (int(v, 0) for v in shp.split(":"))
        capi.set_param("duo", 1 if ilp else -1); capi.set_param("ilp", ilp); capi.set_param("minb", minb); capi.set_param("duo_mask", mask)
        for rng in (0, 120, 60, 30):
            if rng and rng >= b.info().num_trees:
                continue
            capi.set_param("range_trees", rng)
            print(json.dumps(dict(model=m, shape=shp, range_trees=rng, ms=round(t(b), 4), kernel=capi.last_predict_kernel())), flush=True)
    capi.set_param("range_trees", 0); capi.set_param("duo", -1); capi.set_param("ilp", 0); capi.set_param("minb", 0); capi.set_param("duo_mask", 0)
    b.free()
