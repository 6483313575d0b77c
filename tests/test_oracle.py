"""CPU tests of the oracle (test infrastructure) — PARITY UNPINNED: the reference ships no golden
vectors and libxgboost 1.6.0 is absent.  What pins the oracle instead:
  * hand-derivable known-answer boosters (expected values written out below)
  * the committed golden vectors in tests/golden/ made by the pure-Python / numpy restatements
    (tools/make_golden.py), which share no code with oracle/qc_oracle.c
  * agreement of the three model-file readers on the same forest
"""
import os

import numpy as np
import pytest

from oracle import naive, naive_run1
from quickchem_b200 import synth, xgbmodel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _write(forest, tmp_path, name="m.model", **kw):
    p = str(tmp_path / name)
    xgbmodel.write_legacy_binary(forest, p, **kw)
    return p


def test_stump_known_answer(oracle, tmp_path):
    f = xgbmodel.Forest(trees=[xgbmodel.tree_from_nested((3, 0.5, True, -1.0, 2.0))], base_score=0.5, num_feature=27)
    x = np.zeros((6, 27), np.float32)
    x[:, 3] = [0.0, 0.5, np.nextafter(np.float32(0.5), np.float32(0)), 1.0, -999.0, np.nan]
    m = oracle.Model(_write(f, tmp_path))
    assert np.array_equal(m.predict(x), np.float32(0.5) + np.array([-1, 2, -1, 2, -1, -1], np.float32))
    assert np.array_equal(m.predict(x, option_mask=2)[:, 0], np.array([1, 2, 1, 2, 1, 1], np.float32))
    # default right instead
    f.trees[0].default_left[0] = 0
    m = oracle.Model(_write(f, tmp_path, "r.model"))
    assert np.array_equal(m.predict(x), np.float32(0.5) + np.array([-1, 2, -1, 2, 2, 2], np.float32))


def test_sum_order_known_answer(oracle, tmp_path):
    vals = [1e8, 1.0, -1e8, 1.0] * 25
    f = xgbmodel.Forest(trees=[xgbmodel.tree_from_nested(float(v)) for v in vals], base_score=0.5, num_feature=27)
    acc = np.float32(0.5)
    for v in vals:
        acc = np.float32(acc + np.float32(v))
    got = oracle.Model(_write(f, tmp_path)).predict(np.zeros((2, 27), np.float32))
    assert np.all(got == acc)


def test_zero_is_not_missing_and_inf_rejected(oracle, tmp_path):
    f = xgbmodel.Forest(trees=[xgbmodel.tree_from_nested((0, 0.0, True, 1.0, 2.0))], base_score=0.0, num_feature=27)
    m = oracle.Model(_write(f, tmp_path))
    x = np.zeros((2, 27), np.float32)
    assert np.all(m.predict(x) == 2.0)  # 0.0 < 0.0 is false -> right; 0.0 is a value, not missing
    assert np.all(m.predict(x, missing=0.0) == 1.0)  # unless the caller says so
    x[0, 5] = np.inf
    with pytest.raises(oracle.OracleError, match="inf"):
        m.predict(x)
    assert m.predict(x, missing=np.inf).shape == (2,)


def test_golden_predict_all_readers(oracle):
    g = np.load(os.path.join(GOLD, "tiny_predict.npz"))
    x, sums, leaves = g["x"], g["sums"], g["leaves"]
    # C oracle on the two legacy files
    for name in ("tiny.model", "tiny_nobinf.bin"):
        m = oracle.Model(os.path.join(GOLD, name))
        assert np.array_equal(m.predict(x).view(np.uint32), sums.view(np.uint32))
        assert np.array_equal(m.predict(x, option_mask=2).astype(np.int32), leaves)
    # numpy traverser on all three formats
    for rd, name in ((naive.read_legacy, "tiny.model"), (naive.read_json, "tiny.json"), (naive.read_ubj, "tiny.ubj")):
        nm = rd(os.path.join(GOLD, name))
        assert np.array_equal(naive.leaf_ids(nm, x).astype(np.int32), leaves), name
        assert np.array_equal(naive.predict(nm, x).view(np.uint32), sums.view(np.uint32)), name


def test_oracle_vs_naive_random_forest(oracle, tmp_path):
    f = synth.random_forest_structure(25, 9, seed=4)
    rng = np.random.default_rng(5)
    x = rng.normal(0, 1, (3000, 27)).astype(np.float32)
    x[rng.random(x.shape) < 0.05] = -999.0
    p = _write(f, tmp_path)
    m, nm = oracle.Model(p), naive.read_legacy(p)
    assert np.array_equal(m.predict(x, option_mask=2).astype(np.int64), naive.leaf_ids(nm, x))
    assert np.array_equal(m.predict(x).view(np.uint32), naive.predict(nm, x).view(np.uint32))
    assert np.array_equal(m.predict(x, ntree_limit=7).view(np.uint32), naive.predict(nm, x, ntree_limit=7).view(np.uint32))


def test_julian_day(oracle):
    for nymd, want in ((20220101, 1), (20220301, 60), (20240301, 61), (20241231, 366), (19000301, 60), (20000301, 61)):
        assert oracle.julian_day(nymd) == want == naive_run1.julian_day(nymd)


def test_noon_sza_is_about_lat_minus_declination(oracle):
    lat, lon = synth.cubed_sphere_latlon(8)
    sza = oracle.noon_sza(172, lat, lon, synth.MAPL["RADIANS_TO_DEGREES"], synth.MAPL["DEGREES_TO_RADIANS"])
    approx = synth.noon_sza_deg(172, lat)
    assert np.max(np.abs(sza - approx)) < 0.05  # loct ~ 0 (mod 2 pi)
    assert sza.min() >= 0 and sza.max() <= 180


def test_golden_run1(oracle):
    g = np.load(os.path.join(GOLD, "run1_c2.npz"))
    fields = {k[2:]: g[k] for k in g.files if k.startswith("f_")}
    m = oracle.Model(os.path.join(GOLD, "tiny.model"))
    r = oracle.run1(m, fields, synth.MAPL, nymd=int(g["nymd"]), want_features=True)
    assert r["k1"] == int(g["k1"])
    assert np.array_equal(r["X"].view(np.uint32), g["X"].view(np.uint32))
    assert np.array_equal(r["pred"].view(np.uint32), g["pred"].view(np.uint32))
    assert np.array_equal(r["NDWET"].view(np.uint32), g["NDWET"].view(np.uint32))
    # 10**x: powf (C oracle) vs float64 pow rounded once (golden): <= 1 ulp apart
    rel = np.abs(r["OH"].astype(np.float64) - g["OH"]) / np.maximum(np.abs(g["OH"]), 1e-300)
    assert rel.max() <= 1.3e-7


def test_run1_vs_numpy_restatement(oracle, small_model_path):
    """C oracle vs the numpy restatement on C6 x 72 (the DN sums restart at every level)."""
    fields = synth.raw_fields(6)
    m = oracle.Model(small_model_path)
    r = oracle.run1(m, fields, synth.MAPL, nymd=20220726, want_features=True, compute_once_per_day=False)
    sza = oracle.noon_sza(naive_run1.julian_day(20220726), fields["LATS"], fields["LONS"],
                          synth.MAPL["RADIANS_TO_DEGREES"], synth.MAPL["DEGREES_TO_RADIANS"])  # fmt: skip
    feats = naive_run1.features(fields, synth.MAPL, 20220726, sza)
    k1 = naive_run1.level_slab(feats[1], fields["TROPP"], False, 4000.0)
    assert r["k1"] == k1
    X = naive_run1.pack(feats, k1)
    assert np.array_equal(r["X"].view(np.uint32), X.view(np.uint32))
    # a suffix scan would NOT reproduce the DN features (SURVEY.md hard part 6)
    suffix = np.cumsum(fields["TAUCLW"][::-1], axis=0, dtype=np.float32)[::-1]
    assert not np.array_equal(suffix, feats[15])


def test_model_rejects(oracle, tmp_path):
    f = synth.random_forest_structure(2, 3, seed=1)
    p = _write(f, tmp_path, objective="binary:logistic")
    with pytest.raises(oracle.OracleError, match="objective"):
        oracle.Model(p)
    raw = open(_write(f, tmp_path, "ok.model"), "rb").read()
    open(tmp_path / "trunc.model", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(oracle.OracleError, match="truncated"):
        oracle.Model(str(tmp_path / "trunc.model"))
