"""Shared fixtures.  `-m "not gpu"` runs here on CPU (oracle, loaders, host logic, symbol export);
`-m gpu` are the parity tests proper: CUDA path (through the C ABI) vs the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def capi():
    from quickchem_b200 import capi as m

    m.lib()
    return m


@pytest.fixture(params=[0, 1], ids=["nodes8", "duo"])
def duo_mode(request, capi):
    """Run a GPU test twice: on the 8-byte depth-ordered nodes only (duo=0), and with the two-level records
    forced on for every clean-matrix launch (duo=1).  Both must match the oracle bit for bit."""
    capi.set_param("duo", request.param)
    yield request.param
    capi.set_param("duo", -1)


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu

    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def small_forest():
    """27-feature booster grown on C12 synthetic fields: 12 trees, depth <= 8."""
    from quickchem_b200 import synth

    return synth.prod_like_booster(n_trees=12, max_depth=8, n_sample=20000, grid_n=12, seed=5)


@pytest.fixture(scope="session")
def small_model_path(small_forest, tmp_path_factory):
    from quickchem_b200 import xgbmodel

    p = tmp_path_factory.mktemp("models") / "small.model"
    xgbmodel.write_legacy_binary(small_forest, str(p))
    return str(p)


def inject_specials(x, forest, rng, frac_missing=0.01, frac_nan=0.005, frac_on_threshold=0.01):
    """-999.0 / NaN entries and values placed exactly on split thresholds."""
    x = x.copy()
    n, nf = x.shape
    m = rng.random(x.shape)
    x[m < frac_missing] = np.float32(-999.0)
    x[(m >= frac_missing) & (m < frac_missing + frac_nan)] = np.nan
    thr = {}
    for t in forest.trees:
        internal = t.left != -1
        for f, c in zip(t.split_index[internal], t.split_cond[internal]):
            thr.setdefault(int(f), []).append(c)
    k = int(frac_on_threshold * n)
    for f, vals in thr.items():
        rows = rng.integers(0, n, k)
        x[rows, f] = rng.choice(np.asarray(vals, np.float32), k)
    return x
