"""Malformed model files must be rejected (or accepted) — never crash the host process.  A bounded slice of
tools/fuzz_loader.py: truncations, byte flips, insertions and length-field overwrites of the three formats."""
import numpy as np
import pytest

from quickchem_b200 import synth, xgbmodel


@pytest.mark.parametrize("ext,writer", [("model", xgbmodel.write_legacy_binary), ("json", xgbmodel.write_json),
                                        ("ubj", xgbmodel.write_ubj)])  # fmt: skip
def test_mutated_model_files_do_not_crash(capi, tmp_path, ext, writer):
    f = synth.random_forest_structure(3, 4, seed=2)
    src = tmp_path / ("m." + ext)
    writer(f, str(src))
    raw = bytearray(src.read_bytes())
    rng = np.random.default_rng(11)
    outcomes = {"ok": 0, "rejected": 0}
    p = tmp_path / ("fz." + ext)
    for _ in range(250):
        b = bytearray(raw)
        mode = rng.integers(4)
        if mode == 0:
            b = b[: rng.integers(0, len(b))]
        elif mode == 1:
            for _ in range(rng.integers(1, 6)):
                b[rng.integers(len(b))] = rng.integers(256)
        elif mode == 2:
            i = rng.integers(len(b))
            b[i:i] = bytes(rng.integers(0, 256, rng.integers(1, 9), dtype=np.uint8))
        else:
            i = rng.integers(len(b) - 8)
            b[i : i + 4] = int(rng.integers(0, 2**31)).to_bytes(4, "little")
        p.write_bytes(bytes(b))
        try:
            bo = capi.Booster(str(p), parse_only=True)
            bo.info(), bo.flat()
            outcomes["ok"] += 1
        except capi.QcohError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 0 and sum(outcomes.values()) == 250
