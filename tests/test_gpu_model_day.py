"""BASELINE.json configs[4] as a host would drive it — the executable stand-in for fortran/OH_Run1_fused.F90
(this box has no Fortran compiler): one model day of 24 hourly Run1 steps at C180 x 72 through the same C entry
points in the same order, fields resident in HBM, with the Run1 control decisions taken by the library's helpers:

  need_to_call_BOOST   qcoh_need_to_call_boost   OH_GridCompMod.F90:1189-1193 (compute_once_per_day: 00:00:00 only)
  spin-up switch       qcoh_use_inst_values      :1307-1317 (ONLINE_AVG24 and T_avg24(1,1,1) == 0)
  import selection     qcoh_import_name          :1326-1436, :1493-1525
  persistent OH_ML     qcoh_oh_run1              :76-78, :1579-1595 (mask / NDWET follow the CURRENT model state)

OH is compared with the oracle's Run1 at step 1 (the boost step) and at step 13 (OH_ML re-used, model state moved)."""
import numpy as np
import pytest

from quickchem_b200 import synth

pytestmark = [pytest.mark.gpu]

SEL = ("T", "Q", "PLE", "ZLE", "TAUCLW", "TAUCLI", "FCLD", "CH4", "CO") + tuple(s + "SCACOEF" for s in synth.SCA_SPECIES)


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("grid,source,spinup", [(180, 3, False), (48, 3, True), (48, 2, False)],
                         ids=["c180-avg24", "c48-avg24-spinup", "c48-inst"])
def test_model_day(capi, oracle, grid, source, spinup):
    import bench

    model_path = bench.booster_path()
    base = synth.raw_fields(grid)
    km, ncol = base["T"].shape
    # the import state: online 'X', daily means 'X_avg24' (distinct T and PLE, so that boost state != model state)
    imports = dict(base)
    for f in SEL:
        imports[f + "_avg24"] = base[f]
    imports["T_avg24"] = (base["T"] + np.float32(1.5)).astype(np.float32)
    imports["PLE_avg24"] = (base["PLE"] * np.float32(0.999)).astype(np.float32)
    if spinup:  # the coupler hands an all-zero field during the first 24 hours (:1313-1316)
        imports["T_avg24"] = np.zeros_like(base["T"])
    use_inst = bool(capi.lib().qcoh_use_inst_values(source, float(imports["T_avg24"].flat[0])))
    assert use_inst == (source == 3 and spinup)
    chosen = {}
    for f in SEL:
        name, four_d = capi.import_name(f, source, use_inst)
        assert four_d == (f.endswith("SCACOEF"))  # online scattering coefficients carry a wavelength axis
        chosen[f] = imports[name]
    expect_avg = source == 3 and not spinup
    assert (chosen["T"] is imports["T_avg24"]) == expect_avg and (chosen["PLE"] is imports["PLE_avg24"]) == expect_avg
    boost_fields = dict(base)
    boost_fields.update(chosen)

    def model_state(hour):  # the current model state drifts through the day
        m = dict(base)
        m["T"] = (base["T"] + np.float32(0.05 * hour)).astype(np.float32)
        m["PLE"] = (base["PLE"] * np.float32(1.0 + 2e-5 * hour)).astype(np.float32)
        return m

    dev = {k: capi.DeviceArray(v) for k, v in boost_fields.items()}  # resident in HBM for the whole day
    b = capi.Booster(model_path)
    oh = capi.OhRun1(b, ncol, km, synth.MAPL, compute_once_per_day=True)
    om = oracle.Model(model_path)
    oracle.use_all_cores()
    n_predict = capi.kernel_launches("soa_duo")
    k1_day = None
    for hour in range(24):
        nhms = hour * 10000
        need = bool(capi.lib().qcoh_need_to_call_boost(1, nhms))
        assert need == (hour == 0)
        mod = model_state(hour)
        mdev = {k: (capi.DeviceArray(mod[k]) if k in ("T", "PLE") else dev[k]) for k in ("T", "Q", "PLE", "TROPP")}
        rin = oh.make_in(dev, nymd=20220715, need_to_call_boost=need, mod_fields=mdev)
        want = ("OH", "OH_boost") if hour in (0, 12) else ("OH",)
        got = oh.run(rin, want=want)
        if hour == 0:
            k1_day = got["k1"]
            assert 1 < k1_day < 72
        if hour in (0, 12):
            ref = oracle.run1(om, boost_fields, synth.MAPL, nymd=20220715, mod_fields=mod)
            assert _rel(got["OH"], ref["OH"]) <= 1e-6, hour
            if hour == 0:
                assert got["k1"] == ref["k1"]
                assert _rel(got["OH_boost"], ref["OH_boost"]) <= 1e-6
                oh_boost_0 = got["OH_boost"]
            else:  # OH_boost still holds the boost step's values: nothing was predicted since
                assert np.array_equal(got["OH_boost"], oh_boost_0)
    assert capi.kernel_launches("soa_duo") == n_predict + 1  # one prediction per day
    # the DIAG_PL export is bb%PL: built from the PLE handed to boost, not from the model state (:1488, :1666)
    pl_bst = ((boost_fields["PLE"][:-1] + boost_fields["PLE"][1:]) * np.float32(0.5)).astype(np.float32)
    assert np.array_equal(oh.get_diag("PL"), pl_bst)
    oh.free()
