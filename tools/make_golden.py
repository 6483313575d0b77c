#!/usr/bin/env python
"""Generate tests/golden/* (committed).  The reference ships no golden vectors (SURVEY.md §4, §8c:
parity unpinned), so these are produced by the deliberately naive pure-Python / numpy
restatements (oracle/naive.py `predict_scalar`, oracle/naive_run1.py) — an implementation that
shares no code with oracle/qc_oracle.c or the CUDA kernels — and every implementation is then
tested against them.  Run from the repo root:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu, naive, naive_run1  # noqa: E402
from quickchem_b200 import synth, xgbmodel  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

# ---- 1. tiny booster in the three on-disk formats + predictions by the pure-Python walker
forest = synth.random_forest_structure(6, 5, seed=2022, p_leaf=0.2, thr_scale=1.0, base_score=0.5)
xgbmodel.write_legacy_binary(forest, os.path.join(OUT, "tiny.model"))
xgbmodel.write_legacy_binary(forest, os.path.join(OUT, "tiny_nobinf.bin"), with_binf=False, objective="reg:linear")
xgbmodel.write_json(forest, os.path.join(OUT, "tiny.json"))
xgbmodel.write_ubj(forest, os.path.join(OUT, "tiny.ubj"))
rng = np.random.default_rng(726)
x = rng.normal(0, 1, (160, 27)).astype(np.float32)
x[rng.random(x.shape) < 0.03] = np.float32(-999.0)
x[rng.random(x.shape) < 0.02] = np.nan
thr = [(int(f), c) for t in forest.trees for f, c, l in zip(t.split_index, t.split_cond, t.left) if l != -1]
for i, (f, c) in enumerate(thr[:60]):
    x[i, f] = c  # exactly on a threshold -> goes right
x[150:, :] = 0.0
x[151, 0] = -0.0
nm = naive.read_legacy(os.path.join(OUT, "tiny.model"))
sums, leaves = [], []
for r in x:
    s, ids = naive.predict_scalar(nm, r)
    sums.append(s), leaves.append(ids)
np.savez_compressed(os.path.join(OUT, "tiny_predict.npz"), x=x, sums=np.asarray(sums, np.float32),
                    leaves=np.asarray(leaves, np.int32))

# ---- 2. Run1 on a C2 x 12-level grid: features / slab / OH by the numpy restatement
fields = synth.raw_fields(2, seed=11, km=12)
nymd = 20240301
jday = naive_run1.julian_day(nymd)
sza = cpu.noon_sza(jday, fields["LATS"], fields["LONS"], synth.MAPL["RADIANS_TO_DEGREES"], synth.MAPL["DEGREES_TO_RADIANS"])
feats = naive_run1.features(fields, synth.MAPL, nymd, sza)
pl_mod = feats[1]
k1 = naive_run1.level_slab(pl_mod, fields["TROPP"], True, 4000.0)
X = naive_run1.pack(feats, k1)
pred = naive.predict(nm, X)
oh, oh_ml, ndwet = naive_run1.export(pred, k1, fields, synth.MAPL, 0.85)
np.savez_compressed(os.path.join(OUT, "run1_c2.npz"), nymd=nymd, jday=jday, k1=k1, X=X, pred=pred, OH=oh,
                    OH_boost=oh_ml, NDWET=ndwet, sza=sza, **{"f_" + k: v for k, v in fields.items()})
print("golden written to", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})
