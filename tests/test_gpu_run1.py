"""GPU parity of the fused Run1 path (qcoh_oh_run1) and of the host mirror of
`predict_OH_with_XGB` against the oracle's Run1 restatement (OH_GridCompMod.F90:1232-1599).

Bars: the assembled feature matrix X (all 27 features, incl. the O(km^2) restarted vertical sums
and the host-libm noon SZA), the level slab k1, and the raw booster output are bit-exact; OH,
OH_boost within 1e-6 relative (10**x: float64 exp10 on device vs libm powf); NDWET bit-exact."""
import numpy as np
import pytest

from quickchem_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("duo_mode")]
REL_TOL_OH = 1e-6


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    d = np.abs(a - b)
    return float(np.max(np.where(d == 0, 0.0, d / np.maximum(np.abs(b), 1e-300))))


def _run_both(capi, oracle, model_path, fields, **kw):
    km, ncol = fields["T"].shape
    b = capi.Booster(model_path)
    oh = capi.OhRun1(b, ncol, km, synth.MAPL, ohscale=kw.get("ohscale", 0.85),
                     compute_once_per_day=kw.get("compute_once_per_day", True))  # fmt: skip
    rin = oh.make_in(fields, nymd=kw.get("nymd", 20220701), mod_fields=kw.get("mod_fields"))
    got = oh.run(rin, want=("OH", "OH_boost", "NDWET", "X", "pred"))
    ref = oracle.run1(oracle.Model(model_path), fields, synth.MAPL, ohscale=kw.get("ohscale", 0.85),
                      compute_once_per_day=kw.get("compute_once_per_day", True), nymd=kw.get("nymd", 20220701),
                      mod_fields=kw.get("mod_fields"), want_features=True)  # fmt: skip
    return got, ref, oh, rin


def _assert_parity(got, ref):
    assert got["k1"] == ref["k1"]
    assert got["X"].shape == ref["X"].shape
    for f in range(27):
        assert np.array_equal(got["X"][:, f].view(np.uint32), ref["X"][:, f].view(np.uint32)), synth.FEATURE_NAMES[f]
    assert np.array_equal(got["pred"].view(np.uint32), ref["pred"].view(np.uint32))
    assert np.array_equal(got["NDWET"].view(np.uint32), ref["NDWET"].view(np.uint32))
    assert _rel(got["OH_boost"], ref["OH_boost"]) <= REL_TOL_OH
    assert _rel(got["OH"], ref["OH"]) <= REL_TOL_OH


@pytest.mark.parametrize("n", [6, 24])
def test_run1_once_per_day(capi, oracle, small_model_path, n):
    """configs[0]: C24 x 72 single-timestep OH prediction (and a smaller grid)."""
    fields = synth.raw_fields(n)
    got, ref, _, _ = _run_both(capi, oracle, small_model_path, fields)
    assert 1 < got["k1"] < 72  # the 40 hPa slab, not all levels
    _assert_parity(got, ref)
    # levels above the slab are zero in OH_boost (self%OH_ML = 0.0, :1559)
    assert np.all(got["OH_boost"][: got["k1"] - 1] == 0)


def test_run1_fused_path_equals_matrix_path(capi, oracle, small_model_path):
    """Without `X` in the request qcoh_oh_run1 never forms the [N x 27] matrix (predict_soa_kernel reads the
    SoA fields); the outputs must be the very same bits as the matrix path, with and without missing data."""
    for poke in (False, True):
        fields = dict(synth.raw_fields(8))
        if poke:
            for name, where in (("oh_NO2", (40, 7)), ("T", (60, 100)), ("oh_ALBUV", (33,))):
                fields[name] = fields[name].copy()
                fields[name][where] = -999.0
            fields["CO"] = fields["CO"].copy()
            fields["CO"][50, 3] = np.nan
        km, ncol = fields["T"].shape
        oh = capi.OhRun1(capi.Booster(small_model_path), ncol, km, synth.MAPL)
        a = oh.run(oh.make_in(fields), want=("OH", "OH_boost", "X", "pred"))
        b = oh.run(oh.make_in(fields), want=("OH", "OH_boost", "pred"))
        c = oh.run(oh.make_in(fields), want=("OH", "OH_boost"))
        ref = oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL, want_features=True)
        assert np.array_equal(a["pred"].view(np.uint32), ref["pred"].view(np.uint32))
        for k in ("OH", "OH_boost", "pred"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (poke, k)
        assert np.array_equal(a["OH"].view(np.uint32), c["OH"].view(np.uint32))
    fields["oh_O3"] = fields["oh_O3"].copy()
    fields["oh_O3"][70, 5] = np.inf
    with pytest.raises(capi.QcohError, match="inf"):
        oh.run(oh.make_in(fields), want=("OH",))
    with pytest.raises(capi.QcohError, match="inf"):
        oh.run(oh.make_in(fields), want=("OH", "X"))


def test_diag_exports(capi, oracle, small_model_path):
    """DIAG_* exports (OH_GridCompMod.F90:1602-1735): the derived fields of the last boost step, full
    [km][ncol] (not only the predicted slab), bit-exact against the oracle's Run1."""
    fields = synth.raw_fields(6)
    km, ncol = fields["T"].shape
    oh = capi.OhRun1(capi.Booster(small_model_path), ncol, km, synth.MAPL)
    with pytest.raises(capi.QcohError, match="no boost step"):
        oh.get_diag("AODUP")
    got = oh.run(oh.make_in(fields), want=("OH", "OH_boost", "NDWET"))
    ref = oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL, want_features=True)
    for name, f in (("TAUCLWDN", 15), ("TAUCLIDN", 16), ("TAUCLIUP", 17), ("TAUCLWUP", 18), ("AODUP", 23), ("AODDN", 24),
                    ("LAT", 0), ("stratO3", 21), ("SZA", 26), ("PL", 1)):  # fmt: skip
        assert np.array_equal(oh.get_diag(name).view(np.uint32), ref["feat"][f].view(np.uint32)), name
    assert np.array_equal(oh.get_diag("NDWET"), got["NDWET"]) and np.array_equal(oh.get_diag("OH_boost"), got["OH_boost"])
    with pytest.raises(capi.QcohError, match="unknown"):
        oh.get_diag("nope")
    # the noon-SZA cache follows the CONTENTS of LATS / LONS, not only their addresses (a host that re-uses its buffers)
    sza0 = oh.get_diag("SZA")
    f2 = dict(fields)
    f2["LATS"] = fields["LATS"]
    fields["LATS"][:] = (fields["LATS"] * np.float32(0.5)).astype(np.float32)  # same buffer, new grid
    oh.run(oh.make_in(f2), want=("OH",))
    ref2 = oracle.run1(oracle.Model(small_model_path), f2, synth.MAPL, want_features=True)
    assert not np.array_equal(oh.get_diag("SZA"), sza0)
    assert np.array_equal(oh.get_diag("SZA").view(np.uint32), ref2["feat"][26].view(np.uint32))


def test_run1_dynamic_k_range(capi, oracle, small_model_path):
    fields = synth.raw_fields(8)
    got, ref, _, _ = _run_both(capi, oracle, small_model_path, fields, compute_once_per_day=False, nymd=20240229)
    _assert_parity(got, ref)


def test_run1_distinct_model_state(capi, oracle, small_model_path):
    """ONLINE_AVG24 / PRECOMPUTED: model-state T/Q/PLE differ from the boost-state ones."""
    fields = synth.raw_fields(8)
    mod = synth.raw_fields(8, seed=99)
    got, ref, oh, _ = _run_both(capi, oracle, small_model_path, fields, mod_fields=mod)
    _assert_parity(got, ref)
    # DIAG_PL is bb%PL = PL_BST (from the PLE handed to boost, :1488,:1666), not the model state's PL_MOD; DIAG_AOD (:1690)
    pl_bst = ((fields["PLE"][:-1] + fields["PLE"][1:]) * np.float32(0.5)).astype(np.float32)
    pl_mod = ((mod["PLE"][:-1] + mod["PLE"][1:]) * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(oh.get_diag("PL"), pl_bst) and np.array_equal(oh.get_diag("PL_MOD"), pl_mod)
    assert not np.array_equal(pl_bst, pl_mod)
    assert np.array_equal(oh.get_diag("PL").reshape(-1) / np.float32(100.0), ref["X"][:, 1]) or got["k1"] > 1
    sca = sum(fields[s + "SCACOEF"] for s in synth.SCA_SPECIES[1:])
    assert oh.get_diag("AOD").shape == fields["T"].shape
    assert np.array_equal(np.cumsum(oh.get_diag("AOD"), axis=0, dtype=np.float32)[0], oh.get_diag("AODUP")[0]) and sca.shape == fields["T"].shape


def test_run1_tropopause_assert(capi, oracle, small_model_path):
    fields = dict(synth.raw_fields(4))
    fields["TROPP"] = fields["TROPP"].copy()
    fields["TROPP"][5] = 3999.0
    b = capi.Booster(small_model_path)
    oh = capi.OhRun1(b, fields["T"].shape[1], 72, synth.MAPL)
    with pytest.raises(capi.QcohError, match="tropopause"):
        oh.run(oh.make_in(fields))
    with pytest.raises(oracle.OracleError, match="tropopause"):
        oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL)


def test_run1_persistent_oh_ml_and_device_fields(capi, oracle, small_model_path):
    """configs[4] semantics: fields resident in HBM, boost once, later steps reuse OH_ML while
    PL / TROPP / NDWET follow the current model state (OH_GridCompMod.F90:1189-1193,1579-1595)."""
    fields = synth.raw_fields(8)
    km, ncol = fields["T"].shape
    dev = {k: capi.DeviceArray(v) for k, v in fields.items()}
    b = capi.Booster(small_model_path)
    oh = capi.OhRun1(b, ncol, km, synth.MAPL)
    first = oh.run(oh.make_in(dev, need_to_call_boost=True))
    ref = oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL)
    assert _rel(first["OH"], ref["OH"]) <= REL_TOL_OH
    # an hour later: new model state, boost skipped
    later = synth.raw_fields(8, seed=4242)
    step2 = dict(dev)
    for k in ("T", "Q", "PLE", "TROPP"):
        step2[k] = capi.DeviceArray(later[k])
    got = oh.run(oh.make_in(fields=dev, mod_fields=step2, need_to_call_boost=False))
    # expected: same OH_ML, masked / converted with the new state
    pl = (later["PLE"][:-1] + later["PLE"][1:]) * np.float32(0.5)
    tv = later["T"] * (np.float32(1.0) + later["Q"] / synth.MAPL["EPSILON"]) / (np.float32(1.0) + later["Q"])
    ndwet = (synth.MAPL["AVOGAD"] * pl) / (synth.MAPL["RUNIV"] * tv)
    oh_sel = np.where(pl > later["TROPP"][None, :], first["OH_boost"], fields["oh_OH"])
    expect = (oh_sel * ndwet) * np.float32(1.0e-6)
    assert np.array_equal(got["OH"].view(np.uint32), expect.astype(np.float32).view(np.uint32))


def test_run1_rejects_a_malformed_date(capi, small_model_path):
    fields = synth.raw_fields(4)
    oh = capi.OhRun1(capi.Booster(small_model_path), fields["T"].shape[1], 72, synth.MAPL)
    for nymd in (20221301, 20220230, 20220000, -5):
        with pytest.raises(capi.QcohError, match="yyyymmdd"):
            oh.run(oh.make_in(fields, nymd=nymd))
    oh.run(oh.make_in(fields, nymd=20240229))  # a leap day is a date


def test_run1_needs_a_boost_call_first(capi, small_model_path):
    fields = synth.raw_fields(4)
    oh = capi.OhRun1(capi.Booster(small_model_path), fields["T"].shape[1], 72, synth.MAPL)
    with pytest.raises(capi.QcohError, match="need_to_call_boost"):
        oh.run(oh.make_in(fields, need_to_call_boost=False))


def test_diag_partial_sums(capi, small_model_path):
    """Build-defined diagnostic (not in the reference): float64 numpy is the oracle, 1e-10 rel."""
    fields = synth.raw_fields(8)
    km, ncol = fields["T"].shape
    area = (np.random.default_rng(1).random(ncol) * 1e9 + 1e9).astype(np.float32)
    oh = capi.OhRun1(capi.Booster(small_model_path), ncol, km, synth.MAPL)
    got = oh.run(oh.make_in(fields, area=area), want=("OH", "NDWET"))
    f64 = lambda a: a.astype(np.float64)
    pl = (fields["PLE"][:-1] + fields["PLE"][1:]) * np.float32(0.5)
    trop = pl > fields["TROPP"][None, :]
    dp = f64(fields["PLE"][1:]) - f64(fields["PLE"][:-1])
    dz = f64(fields["ZLE"][:-1]) - f64(fields["ZLE"][1:])
    w = dp * f64(area)[None, :] / 9.80665
    nch4 = f64(fields["CH4"]) * f64(got["NDWET"])
    kt = 2.45e-12 * np.exp(-1775.0 / f64(fields["T"]))
    vol = f64(area)[None, :] * dz
    expect = np.array([(f64(got["OH"]) * w)[trop].sum(), w[trop].sum(), (nch4 * vol)[trop].sum(),
                       (kt * f64(got["OH"]) * nch4 * vol)[trop].sum()])  # fmt: skip
    assert np.allclose(got["diag"], expect, rtol=1e-10, atol=0)


def test_host_mirror_predict_OH_with_XGB(capi, oracle, small_model_path):
    """The reference's own driver routine, restated on the host and linked against libqcoh's
    XGBoost-named symbols only, against the oracle's Run1 (same inputs)."""
    fields = synth.raw_fields(6)
    km, ncol = fields["T"].shape
    ref = oracle.run1(oracle.Model(small_model_path), fields, synth.MAPL, ohscale=1.0, want_features=True)
    # bb from the oracle-assembled features (feature 2 is PL in Pa before the /100 of :314)
    k1 = ref["k1"]
    pl_mod = ((fields["PLE"][:-1] + fields["PLE"][1:]) * np.float32(0.5)).astype(np.float32)
    X = np.zeros((km * ncol, 27), np.float32)
    X[(k1 - 1) * ncol :] = ref["X"]
    bb = []
    for f in range(27):
        if f in (0, 21, 22, 26):
            bb.append(np.ascontiguousarray(ref["X"][:ncol, f]))
        elif f == 1:
            bb.append(pl_mod)
        else:
            bb.append(np.ascontiguousarray(X[:, f].reshape(km, ncol)))
    capi.lib().qcoh_predict_OH_reset()
    OH_ML = np.zeros((km, ncol), np.float32)
    rc = capi.predict_OH_with_XGB(small_model_path, 6, ncol // 6, km, False, 4000.0, pl_mod, fields["TROPP"], bb, OH_ML)
    assert rc == 0, capi.last_error()
    # dynamic_k_range=False => same 40 hPa slab as the oracle's compute_once_per_day run
    assert np.all(OH_ML[: k1 - 1] == 0)
    expect = np.float32(10.0) ** ref["pred"].reshape(km - k1 + 1, ncol)
    assert _rel(OH_ML[k1 - 1 :], expect.astype(np.float32)) <= REL_TOL_OH
    # second call reuses the SAVEd booster (first_time = .FALSE.)
    OH2 = np.zeros_like(OH_ML)
    assert capi.predict_OH_with_XGB("/nonexistent/ignored.model", 6, ncol // 6, km, False, 4000.0, pl_mod,
                                    fields["TROPP"], bb, OH2) == 0  # fmt: skip
    assert np.array_equal(OH_ML, OH2)
    # the reference's _ASSERT (:287-288) has a message here too, and XGBGetLastError carries it
    low = fields["TROPP"].copy()
    low[3] = 3000.0
    assert capi.predict_OH_with_XGB(small_model_path, 6, ncol // 6, km, False, 4000.0, pl_mod, low, bb, OH2) == -1
    assert "Minimum tropopause pressure is not low enough" in capi.last_error()
    capi.lib().qcoh_predict_OH_reset()
    # a failing INIT (:242-271) reports the loader's reason, frees what it created and can be retried
    for _ in range(3):
        assert capi.predict_OH_with_XGB("/nonexistent/model.bin", 6, ncol // 6, km, False, 4000.0, pl_mod,
                                        fields["TROPP"], bb, OH2) == -1  # fmt: skip
        assert "Opening" in capi.last_error()
    assert capi.predict_OH_with_XGB(small_model_path, 6, ncol // 6, km, False, 4000.0, pl_mod, fields["TROPP"], bb, OH2) == 0
    assert np.array_equal(OH_ML, OH2)
    capi.lib().qcoh_predict_OH_reset()


def test_native_nccl_allreduce_single_rank(capi):
    """qcoh_comm_*: the library's own NCCL communicator (dlopen'ed libnccl).  One GPU here, so one rank: the
    all-reduce must be the identity; bench.py --gpus N exercises N ranks."""
    uid = capi.comm_unique_id()
    assert len(uid) == 128
    capi.comm_init(1, 0, uid)
    try:
        v = np.array([1.5, -2.25, 3e300, 4e-300])
        assert np.array_equal(capi.comm_allreduce_sum(v), v)
        with pytest.raises(capi.QcohError, match="already exists"):
            capi.comm_init(1, 0, uid)
    finally:
        capi.comm_destroy()
    with pytest.raises(capi.QcohError, match="qcoh_comm_init first"):
        capi.comm_allreduce_sum([1.0])



def test_loss_frequencies_for_ch4_co(capi, small_model_path):
    """SURVEY.md 8(f)4 (build-defined, no reference counterpart): LOSS_CH4 = 2.45e-12 exp(-1775/T) [OH] and
    LOSS_CO = 1.5e-13 (1 + 0.6 PL/101325) [OH] from the final OH (molec/cm3), float64 on the device and rounded
    once — checked against numpy float64 on the returned OH, T and DIAG_PL (1e-6 relative, stated here)."""
    fields = synth.raw_fields(6)
    km, ncol = fields["T"].shape
    oh = capi.OhRun1(capi.Booster(small_model_path), ncol, km, synth.MAPL)
    got = oh.run(oh.make_in(fields), want=("OH", "LOSS_CH4", "LOSS_CO"))
    pl = oh.get_diag("PL").astype(np.float64)
    ohn, t = got["OH"].astype(np.float64), fields["T"].astype(np.float64)
    assert _rel(got["LOSS_CH4"], 2.45e-12 * np.exp(-1775.0 / t) * ohn) <= 1e-6
    assert _rel(got["LOSS_CO"], 1.5e-13 * (1.0 + 0.6 * pl / 101325.0) * ohn) <= 1e-6
    assert np.all(got["LOSS_CH4"] > 0) and np.all(got["LOSS_CO"] > got["LOSS_CH4"])
    # not requested => not produced, OH unchanged, also on a non-boost step
    again = oh.run(oh.make_in(fields, need_to_call_boost=False), want=("OH", "LOSS_CO"))
    assert "LOSS_CH4" not in again and np.array_equal(again["OH"], got["OH"])
    assert np.array_equal(again["LOSS_CO"], got["LOSS_CO"])


def _two_month_models(tmp_path):
    from quickchem_b200 import xgbmodel

    paths = {}
    for mm, seed in ((7, 5), (8, 6)):
        f = synth.prod_like_booster(n_trees=6, max_depth=6, n_sample=8000, grid_n=12, seed=seed)
        paths[mm] = str(tmp_path / f"oh_M{mm:02d}.model")
        xgbmodel.write_legacy_binary(f, paths[mm])
    return str(tmp_path / "oh_M%m2.model"), paths


def test_month_rollover_is_opt_in(capi, oracle, tmp_path):
    """the reference expands `..._M%m2.model` every step but predicts with the booster it loaded
    first.  qcoh_oh_run1 keeps that; qcoh_oh_select_model (template + cache + switch) is the opt-in fix."""
    pattern, paths = _two_month_models(tmp_path)
    fields = synth.raw_fields(6)
    km, ncol = fields["T"].shape
    ref = {mm: oracle.run1(oracle.Model(p), fields, synth.MAPL, nymd=20220801) for mm, p in paths.items()}
    assert not np.array_equal(ref[7]["OH_boost"], ref[8]["OH_boost"])
    capi.model_cache_clear()
    july = capi.Booster.cached(capi.expand_template(pattern, 20220731))
    oh = capi.OhRun1(july, ncol, km, synth.MAPL)
    # default: 1 August still runs July's model (reference behaviour)
    got = oh.run(oh.make_in(fields, nymd=20220801), want=("OH", "OH_boost"))
    assert _rel(got["OH_boost"], ref[7]["OH_boost"]) <= REL_TOL_OH
    assert oh.select_model(pattern, 20220731) is False
    assert oh.select_model(pattern, 20220801) is True and capi.model_cache_size() == 2
    got = oh.run(oh.make_in(fields, nymd=20220801), want=("OH", "OH_boost"))
    assert _rel(got["OH_boost"], ref[8]["OH_boost"]) <= REL_TOL_OH
    assert _rel(got["OH"], ref[8]["OH"]) <= REL_TOL_OH
    # back and forth reuses the cached boosters (no reload), and the constant-memory tops follow the switch
    assert oh.select_model(pattern, 20220715, 120000) is True and capi.model_cache_size() == 2
    got = oh.run(oh.make_in(fields, nymd=20220801), want=("OH", "OH_boost"))
    assert _rel(got["OH_boost"], ref[7]["OH_boost"]) <= REL_TOL_OH
    with pytest.raises(capi.QcohError, match="No such file|cannot open|open"):
        oh.select_model(pattern, 20220901)
    oh.free()
    capi.model_cache_clear()
    assert capi.model_cache_size() == 0


def test_booster_reload_in_place(capi, oracle, tmp_path):
    """XGBoosterLoadModel on a booster that already holds a model replaces it (libxgboost semantics); the host
    mirror uses exactly that when qcoh_predict_OH_reload_on_file_change(1) is set, and never otherwise."""
    _, paths = _two_month_models(tmp_path)
    rng = np.random.default_rng(3)
    x = synth.quick_features(synth.raw_fields(6))[rng.integers(0, 6 * 36 * 72, 5000)]
    b = capi.Booster(paths[7])
    m = capi.DMatrix(x)
    p7 = b.predict(m)
    b.load_model(paths[8])
    p8 = b.predict(m)
    assert np.array_equal(p7, oracle.Model(paths[7]).predict(x)) and np.array_equal(p8, oracle.Model(paths[8]).predict(x))
    assert not np.array_equal(p7, p8)
    # the mirror
    fields = synth.raw_fields(6)
    km, ncol = fields["T"].shape
    pl_mod = ((fields["PLE"][:-1] + fields["PLE"][1:]) * np.float32(0.5)).astype(np.float32)
    X = synth.quick_features(fields)
    bb = [np.ascontiguousarray(X[:ncol, f]) if f in (0, 21, 22, 26) else
          (pl_mod if f == 1 else np.ascontiguousarray(X[:, f].reshape(km, ncol))) for f in range(27)]  # fmt: skip

    def call(path):
        out = np.zeros((km, ncol), np.float32)
        assert capi.predict_OH_with_XGB(path, 6, ncol // 6, km, True, 4000.0, pl_mod, fields["TROPP"], bb, out) == 0
        return out

    L = capi.lib()
    L.qcoh_predict_OH_reset()
    try:
        a7, a8_default = call(paths[7]), call(paths[8])
        assert np.array_equal(a7, a8_default)  # SAVE'd booster: the second file name is ignored
        L.qcoh_predict_OH_reload_on_file_change(1)
        a8 = call(paths[8])
        assert not np.array_equal(a7, a8)
        assert np.array_equal(call(paths[7]), a7)
    finally:
        L.qcoh_predict_OH_reload_on_file_change(0)
        L.qcoh_predict_OH_reset()
