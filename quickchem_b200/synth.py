"""Seeded synthetic inputs for the OH path: cubed-sphere met/chem fields and boosters.

Nothing here is on the product path; it only manufactures inputs (there are no model
files, MERRA2-GMI fields or network on the build/GPU boxes — SURVEY.md §8c/§8d).

Array layout matches MAPL/Fortran `(im, jm, km)` column-major == C order `[k][j][i]`;
here the horizontal index is flattened to a *column* index `c = i + N*j` of the
`(N, 6N)` cubed-sphere index space, so 3-D centre fields are `[km][ncol]`, edge fields
(`PLE`, `ZLE`, lower bound 0 — `OH_GridCompMod.F90:1246,1450`) are `[km+1][ncol]`, and
2-D fields are `[ncol]`.  Level 1 (index 0) is the model top.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
from scipy.signal import lfilter

from .xgbmodel import Forest, Tree

KM = 72
NFEAT = 27
SCA_SPECIES = ("BC", "OC", "BR", "DU", "SU", "SS", "NI")
CLIM_GASES = ("NO2", "O3", "ISOP", "ACET", "C2H6", "C3H8", "PRPE", "ALK4", "MP", "H2O2", "CH2O")
# feature order contract, OH_GridCompMod.F90:313-339 (1-based there)
FEATURE_NAMES = (
    "LAT", "PL", "T", "NO2", "O3", "CH4", "CO", "ISOP", "ACET", "C2H6", "C3H8", "PRPE", "ALK4", "MP",
    "H2O2", "TAUCLWDN", "TAUCLIDN", "TAUCLIUP", "TAUCLWUP", "CLOUD", "QV", "GMISTRATO3", "ALBUV",
    "AODUP", "AODDN", "CH2O", "SZA",
)  # fmt: skip
assert len(FEATURE_NAMES) == NFEAT

# MAPL constants are external to the reference (MAPL_Constants); the values used by GEOS.
MAPL = dict(
    EPSILON=np.float32(18.015 / 28.965),
    AVOGAD=np.float32(6.023e26),
    RUNIV=np.float32(8314.47),
    RADIANS_TO_DEGREES=np.float32(180.0 / np.pi),
    DEGREES_TO_RADIANS=np.float32(np.pi / 180.0),
)


def grid_dims(n: int):
    """(im_g, jm_g, ncol) of cubed sphere C<n> in MAPL's index space."""
    return n, 6 * n, 6 * n * n


def cubed_sphere_latlon(n: int):
    """Equiangular gnomonic cell centres, radians, flattened `c = i + n*j`, face = j // n."""
    a = (np.arange(n, dtype=np.float64) + 0.5) / n * (np.pi / 2) - np.pi / 4
    x, y = np.meshgrid(np.tan(a), np.tan(a), indexing="xy")  # x varies with i (fast), y with j
    one = np.ones_like(x)
    faces = [
        (one, x, y), (-x, one, y), (-one, -x, y), (x, -one, y), (-y, x, one), (y, x, -one),
    ]  # fmt: skip
    lat, lon = [], []
    for fx, fy, fz in faces:
        r = np.sqrt(fx * fx + fy * fy + fz * fz)
        lat.append(np.arcsin(fz / r))
        lon.append(np.arctan2(fy, fx))
    lat = np.concatenate([v.reshape(-1) for v in lat]).astype(np.float32)
    lon = np.concatenate([v.reshape(-1) for v in lon])
    lon = np.where(lon < 0, lon + 2 * np.pi, lon).astype(np.float32)  # GEOS longitudes 0..2pi
    return lat, lon


def _edge_sigma(km: int = KM) -> np.ndarray:
    s = (np.arange(km + 1, dtype=np.float64) / km)
    return 0.85 * s**4 + 0.15 * s**1.5


def _ar1(rng, shape, rho=0.95, nrow_len=None) -> np.ndarray:
    """Unit-variance noise, AR(1)-correlated along the fast horizontal index `i`."""
    eps = rng.standard_normal(shape, dtype=np.float32)
    if nrow_len is None or rho == 0.0:
        return eps
    v = eps.reshape(-1, nrow_len)
    out = lfilter([np.sqrt(1 - rho * rho)], [1.0, -rho], v, axis=-1).astype(np.float32)
    return out.reshape(shape)


_LATLON_CACHE: dict = {}


def raw_fields(n: int, seed: int | None = None, km: int = KM, rho: float = 0.95, col0: int = 0,
               ncol: int | None = None) -> Dict[str, np.ndarray]:
    """Synthetic import state for `OH_data_source = ONLINE_INST` on cubed sphere C<n>.

    Distributions follow SURVEY.md §8(d).  `rho=0` gives the uncorrelated worst case.
    `col0`/`ncol` restrict the state to one rank's contiguous column range (a multiple of `n`
    columns, i.e. whole j-rows, so the AR(1) rows stay intact); seed it per rank.
    """
    rng = np.random.default_rng(20220726 + n if seed is None else seed)
    if n not in _LATLON_CACHE:
        _LATLON_CACHE.clear()
        _LATLON_CACHE[n] = cubed_sphere_latlon(n)
    lat, lon = _LATLON_CACHE[n]
    if ncol is None:
        _, _, ncol = grid_dims(n)
        ncol -= col0
    assert ncol % n == 0 and col0 % n == 0, "shards are whole j-rows of the (N, 6N) index space"
    lat, lon = lat[col0 : col0 + ncol], lon[col0 : col0 + ncol]
    f: Dict[str, np.ndarray] = {"LATS": lat, "LONS": lon}

    def noise3(sig_col=0.6, sig_cell=0.8):
        col = _ar1(rng, (ncol,), rho, n)
        cell = _ar1(rng, (km, ncol), rho, n)
        return (sig_col * col[None, :] + sig_cell * cell).astype(np.float32)

    ps = (98500.0 + 3000.0 * _ar1(rng, (ncol,), rho, n)).astype(np.float32)
    sig = _edge_sigma(km).astype(np.float32)
    ple = (np.float32(1.0) + (ps[None, :] - np.float32(1.0)) * sig[:, None]).astype(np.float32)
    f["PLE"] = ple
    pl = 0.5 * (ple[:-1] + ple[1:])
    # height (m) from a scale-height atmosphere, used for lapse rate and decays
    z = (-7400.0 * np.log(pl / ps[None, :])).astype(np.float32)
    t = np.maximum(288.0 - 6.5e-3 * z, 210.0) + np.maximum(z - 20000.0, 0.0) * 1.6e-3
    f["T"] = (t + 3.0 * _ar1(rng, (km, ncol), rho, n)).astype(np.float32)
    q = 0.012 * np.exp(-z / 2500.0) * np.exp(0.7 * noise3())
    f["Q"] = np.clip(q, 1e-7, 0.03).astype(np.float32)
    # hydrostatic edge heights, top-down index, ZLE(km) = surface = 0
    dz = (287.0 * f["T"] / 9.80665 * np.log(ple[1:] / np.maximum(ple[:-1], np.float32(1.0)))).astype(np.float32)
    zle = np.zeros((km + 1, ncol), np.float32)
    zle[:-1] = np.cumsum(dz[::-1], axis=0)[::-1]
    f["ZLE"] = zle

    def lognormal(mean_k, sigma=0.7):
        m = np.asarray(mean_k, np.float32)
        m = m[:, None] if m.ndim == 1 else m
        return (m * np.exp(sigma * noise3())).astype(np.float32)

    zk = z.mean(axis=1)
    ones = np.ones(km, np.float32)
    f["CH4"] = lognormal(1.8e-6 * ones, 0.05)
    f["CO"] = lognormal(1.0e-7 * ones)
    means = dict(NO2=1e-10, ISOP=1e-10, ACET=5e-10, C2H6=1e-9, C3H8=3e-10, PRPE=5e-11, ALK4=2e-10,
                 MP=3e-10, H2O2=1e-9, CH2O=3e-10)  # fmt: skip
    for g in CLIM_GASES:
        if g == "O3":
            prof = 4e-8 + (5e-6 - 4e-8) / (1.0 + np.exp(-(zk - 22000.0) / 3000.0))
            f["oh_O3"] = lognormal(prof.astype(np.float32), 0.4)
        elif g == "ISOP":
            f["oh_ISOP"] = lognormal((means[g] * np.exp(-zk / 1500.0)).astype(np.float32))
        else:
            f["oh_" + g] = lognormal(means[g] * ones)
    f["oh_OH"] = lognormal(1e-13 * ones)

    in_cloud_band = ((pl > 30000.0) & (pl < 90000.0))
    for name in ("TAUCLW", "TAUCLI"):
        tau = rng.gamma(0.3, 2.0, size=(km, ncol)).astype(np.float32)
        clear = _ar1(rng, (km, ncol), rho, n) < 0.5244  # P(clear) = 0.7, horizontally coherent
        f[name] = np.where(in_cloud_band & ~clear, tau, np.float32(0.0)).astype(np.float32)
    cloudy = (f["TAUCLW"] + f["TAUCLI"]) > 0
    f["FCLD"] = (rng.random((km, ncol), dtype=np.float32) * cloudy).astype(np.float32)
    for sp in SCA_SPECIES:
        f[sp + "SCACOEF"] = lognormal((1e-6 * np.exp(-zk / 3000.0)).astype(np.float32))

    f["oh_GMITO3"] = (300.0 + 40.0 * _ar1(rng, (ncol,), rho, n)).astype(np.float32)
    f["oh_GMITTO3"] = (35.0 + 8.0 * _ar1(rng, (ncol,), rho, n)).astype(np.float32)
    f["oh_ALBUV"] = np.clip(0.46 + 0.25 * _ar1(rng, (ncol,), rho, n), 0.02, 0.9).astype(np.float32)
    f["TROPP"] = (10000.0 + 20000.0 * np.abs(np.sin(lat.astype(np.float64))) ** 1.5).astype(np.float32)
    return f


def field_block_rows(n: int) -> int:
    """j-rows per independently seeded block of `raw_fields_blocked`: 72 blocks per grid when 12 divides n."""
    return n // 12 if n % 12 == 0 else 1


def raw_fields_blocked(n: int, seed: int, j0: int = 0, j1: int | None = None, **kw) -> Dict[str, np.ndarray]:
    """The synthetic state of j-rows [j0, j1) of C<n>, identical whatever the sharding: the (n, 6n) index space is
    cut into blocks of `field_block_rows(n)` whole j-rows, block b is `raw_fields(..., seed = f(seed, b))` (there is
    no correlation across j-rows, so blocks are independent), and a shard is the concatenation of the blocks it
    covers.  1, 2, 4 or 8 ranks therefore see the very same global fields (bench.py checks the results across N)."""
    jm = 6 * n
    j1 = jm if j1 is None else j1
    rows = field_block_rows(n)
    assert j0 % rows == 0 and (j1 % rows == 0 or j1 == jm), "shards are whole blocks of j-rows"
    if n not in _LATLON_CACHE:
        _LATLON_CACHE.clear()
        _LATLON_CACHE[n] = cubed_sphere_latlon(n)

    def block(b):
        r0, r1 = b * rows, min((b + 1) * rows, jm)
        return raw_fields(n, seed=(seed * 1000003 + b) % (2**63), col0=r0 * n, ncol=(r1 - r0) * n, **kw)

    parts = [block(b) for b in range(j0 // rows, (j1 + rows - 1) // rows)]
    return {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=-1)) for k in parts[0]}


def noon_sza_deg(jday: int, lat: np.ndarray) -> np.ndarray:
    """Cheap stand-in for the local-noon SZA (|lat - dec|); only used to make synthetic X."""
    dec = np.arcsin(0.3978 * np.sin(0.9863 * (jday - 80.0) * np.pi / 180.0))
    return np.abs(np.degrees(lat.astype(np.float64)) - np.degrees(dec)).astype(np.float32)


def quick_features(raw: Dict[str, np.ndarray], jday: int = 182) -> np.ndarray:
    """Dense feature matrix `X[N][27]` (row `m = col + ncol*k`, the reference's pack order
    `OH_GridCompMod.F90:308-345` with k1=1) built with fast vectorised numpy.

    This is *input synthesis*, not the reference's assembly: it uses cumulative sums whose
    rounding differs from `SUM(x(:,:,k:km),3)`.  The bit-faithful assembly lives in
    `oracle/` (CPU) and in the CUDA library (K1)."""
    ple, zle = raw["PLE"], raw["ZLE"]
    km, ncol = raw["T"].shape
    pl = (ple[:-1] + ple[1:]) * np.float32(0.5)
    thick = zle[:-1] - zle[1:]
    sca = raw["BCSCACOEF"].copy()
    for sp in SCA_SPECIES[1:]:
        sca += raw[sp + "SCACOEF"]
    aod = thick * sca
    up = lambda a: np.cumsum(a, axis=0, dtype=np.float32)
    dn = lambda a: np.cumsum(a[::-1], axis=0, dtype=np.float32)[::-1]
    two_d = lambda a: np.broadcast_to(a[None, :], (km, ncol))
    cols = [
        two_d(raw["LATS"] * MAPL["RADIANS_TO_DEGREES"]), pl / np.float32(100.0), raw["T"], raw["oh_NO2"],
        raw["oh_O3"], raw["CH4"], raw["CO"], raw["oh_ISOP"], raw["oh_ACET"], raw["oh_C2H6"], raw["oh_C3H8"],
        raw["oh_PRPE"], raw["oh_ALK4"], raw["oh_MP"], raw["oh_H2O2"], dn(raw["TAUCLW"]), dn(raw["TAUCLI"]),
        up(raw["TAUCLI"]), up(raw["TAUCLW"]), raw["FCLD"], raw["Q"],
        two_d(raw["oh_GMITO3"] - raw["oh_GMITTO3"]), two_d(raw["oh_ALBUV"]), up(aod), dn(aod), raw["oh_CH2O"],
        two_d(noon_sza_deg(jday, raw["LATS"])),
    ]  # fmt: skip
    x = np.empty((km * ncol, NFEAT), np.float32)
    for j, c in enumerate(cols):
        x[:, j] = np.ascontiguousarray(c, dtype=np.float32).reshape(-1)
    return x


def synthetic_log10_oh(x: np.ndarray, rng=None) -> np.ndarray:
    """A smooth, OH-like regression target (log10 mol/mol) over the 27 features."""
    f = {n: x[:, i].astype(np.float64) for i, n in enumerate(FEATURE_NAMES)}
    l10 = lambda v, ref: np.log10(np.maximum(v, 1e-30) / ref)
    cosz = np.cos(np.radians(np.minimum(f["SZA"], 89.0)))
    y = (
        -12.9 + 0.9 * np.log10(np.maximum(cosz, 0.02)) + 0.35 * l10(f["O3"], 4e-8) + 0.30 * l10(f["QV"], 1e-3)
        - 0.25 * l10(f["CO"], 1e-7) + 0.20 * l10(f["NO2"], 1e-10) - 0.10 * l10(f["CH4"], 1.8e-6)
        - 0.08 * l10(f["ISOP"] + 1e-12, 1e-10) + 0.05 * l10(f["CH2O"], 3e-10) + 0.04 * l10(f["H2O2"], 1e-9)
        - 0.10 * np.tanh(f["TAUCLWUP"] + f["TAUCLIUP"]) + 0.06 * np.tanh(f["TAUCLWDN"] + f["TAUCLIDN"])
        + 0.15 * (f["ALBUV"] - 0.4) - 0.05 * np.tanh(50.0 * (f["AODUP"] + f["AODDN"]))
        - 0.0008 * (f["GMISTRATO3"] - 265.0) + 0.15 * (f["T"] - 250.0) / 40.0 + 0.1 * l10(f["PL"], 500.0)
    )  # fmt: skip
    if rng is not None:
        y = y + 0.03 * rng.standard_normal(y.shape)
    return y.astype(np.float32)


# --------------------------------------------------------------------------------------
# booster grower (level-wise, best-of-K random candidate splits, squared error)
# --------------------------------------------------------------------------------------
def _grow_tree(x, r, rng, max_depth, min_leaf, n_cand, eta, reg_lambda, default_left_p):
    n, nfeat = x.shape
    left, right, parent, sidx, cond, dl, cover = [-1], [-1], [-1], [0], [0.0], [0], [float(n)]
    depth_of = [0]
    node_of = np.zeros(n, np.int64)  # node id of every still-active sample
    idx = np.arange(n)  # active sample ids
    frontier = np.array([0])
    leaf_sum = {}
    sample_leaf = np.zeros(n, np.int64)
    for depth in range(max_depth + 1):
        if idx.size == 0:
            break
        # local ids of frontier nodes
        lut = np.full(len(left), -1, np.int64)
        lut[frontier] = np.arange(frontier.size)
        loc = lut[node_of[idx]]
        nf = frontier.size
        cnt = np.bincount(loc, minlength=nf).astype(np.float64)
        rs = np.bincount(loc, weights=r[idx], minlength=nf)
        best_gain = np.full(nf, -np.inf)
        best_f = np.zeros(nf, np.int64)
        best_thr = np.zeros(nf, np.float32)
        can_split = (cnt >= 2 * min_leaf) & (depth < max_depth)
        if can_split.any():
            order = np.argsort(loc, kind="stable")
            starts = np.concatenate([[0], np.cumsum(cnt.astype(np.int64))[:-1]])
            base = rs * rs / (cnt + reg_lambda)
            for _ in range(n_cand):
                fk = rng.integers(0, nfeat, nf)
                pick = starts + (rng.random(nf) * cnt).astype(np.int64)
                pick = np.minimum(pick, starts + cnt.astype(np.int64) - 1)
                thr = x[idx[order[pick]], fk]
                key = x[idx, fk[loc]]
                gl = key < thr[loc]
                cl = np.bincount(loc, weights=gl, minlength=nf)
                sl = np.bincount(loc, weights=r[idx] * gl, minlength=nf)
                cr, sr = cnt - cl, rs - sl
                gain = sl * sl / (cl + reg_lambda) + sr * sr / (cr + reg_lambda) - base
                ok = can_split & (cl >= min_leaf) & (cr >= min_leaf)
                better = ok & (gain > best_gain)
                best_gain[better], best_f[better], best_thr[better] = gain[better], fk[better], thr[better]
        do_split = np.isfinite(best_gain)
        # finalise leaves
        for j in np.nonzero(~do_split)[0]:
            nid = frontier[j]
            cond[nid] = float(np.float32(eta * rs[j] / (cnt[j] + reg_lambda)))
        leaf_mask = ~do_split[loc]
        sample_leaf[idx[leaf_mask]] = node_of[idx[leaf_mask]]
        # expand
        new_frontier = []
        child_left = np.zeros(nf, np.int64)
        dls = rng.random(nf) < default_left_p
        for j in np.nonzero(do_split)[0]:
            nid = frontier[j]
            l = len(left)
            for _c in range(2):
                left.append(-1), right.append(-1), parent.append(nid), sidx.append(0), cond.append(0.0)
                dl.append(0), cover.append(0.0), depth_of.append(depth + 1)
            left[nid], right[nid] = l, l + 1
            sidx[nid], cond[nid], dl[nid] = int(best_f[j]), float(best_thr[j]), int(dls[j])
            child_left[j] = l
            new_frontier += [l, l + 1]
        keep = ~leaf_mask
        idx, loc = idx[keep], loc[keep]
        if idx.size:
            go_right = ~(x[idx, best_f[loc]] < best_thr[loc])
            node_of[idx] = child_left[loc] + go_right
            cc = np.bincount(node_of[idx], minlength=len(left))
            for nid in new_frontier:
                cover[nid] = float(cc[nid])
        frontier = np.asarray(new_frontier, np.int64)
    tree = Tree(
        left=np.asarray(left, np.int32), right=np.asarray(right, np.int32), parent=np.asarray(parent, np.int32),
        split_index=np.asarray(sidx, np.uint32), split_cond=np.asarray(cond, np.float32),
        default_left=np.asarray(dl, np.uint8), sum_hess=np.asarray(cover, np.float32),
    )  # fmt: skip
    return tree, sample_leaf


def grow_forest(x, y, n_trees=100, max_depth=18, min_leaf=8, n_cand=3, eta=0.3, base_score=0.5,
                reg_lambda=1.0, default_left_p=0.5, seed=0) -> Forest:  # fmt: skip
    """Gradient-boost `n_trees` regression trees on (x, y); returns an XGBoost-shaped Forest."""
    rng = np.random.default_rng(seed)
    x = np.ascontiguousarray(x, np.float32)
    pred = np.full(x.shape[0], base_score, np.float64)
    trees = []
    for _ in range(n_trees):
        r = y.astype(np.float64) - pred
        tree, sample_leaf = _grow_tree(x, r, rng, max_depth, min_leaf, n_cand, eta, reg_lambda, default_left_p)
        pred += tree.split_cond[sample_leaf].astype(np.float64)
        trees.append(tree)
    return Forest(trees=trees, base_score=base_score, num_feature=x.shape[1])


def random_forest_structure(n_trees, max_depth, num_feature=NFEAT, seed=0, p_leaf=0.15, thr_scale=1.0,
                            base_score=0.5) -> Forest:  # fmt: skip
    """Data-free random forest (random topology/thresholds ~ N(0, thr_scale)) for unit tests."""
    rng = np.random.default_rng(seed)
    trees = []
    for _ in range(n_trees):
        def spec(d):
            if d == max_depth or (d > 0 and rng.random() < p_leaf):
                return float(np.float32(rng.normal(0, 0.3)))
            return (int(rng.integers(num_feature)), float(np.float32(rng.normal(0, thr_scale))),
                    bool(rng.random() < 0.5), spec(d + 1), spec(d + 1))  # fmt: skip
        from .xgbmodel import tree_from_nested
        trees.append(tree_from_nested(spec(0)))
    return Forest(trees=trees, base_score=base_score, num_feature=num_feature)


def replicate_forest(forest: Forest, times: int, jitter_rel: float = 1e-3, seed: int = 0) -> Forest:
    """`times` copies of a grown forest back to back, each copy with its split thresholds jittered by a relative
    N(0, jitter_rel) (so that the copies are distinct in memory and their walks differ a little) — a cheap stand-in for
    a booster with `times` x as many grown trees (booster sweep: 500 and 1000 trees, forests larger than L2)."""
    rng = np.random.default_rng(seed)
    trees = []
    for c in range(times):
        for t in forest.trees:
            cond = t.split_cond.copy()
            if c > 0:
                internal = t.left != -1
                j = (1.0 + jitter_rel * rng.standard_normal(int(internal.sum()))).astype(np.float32)
                cond[internal] = (cond[internal] * j).astype(np.float32)
            trees.append(Tree(left=t.left, right=t.right, parent=t.parent, split_index=t.split_index, split_cond=cond,
                              default_left=t.default_left, sum_hess=t.sum_hess))  # fmt: skip
    return Forest(trees=trees, base_score=forest.base_score, num_feature=forest.num_feature)


def prod_like_booster(n_trees=100, max_depth=18, n_sample=131072, min_leaf=8, seed=18, grid_n=48,
                      n_cand=3) -> Forest:  # fmt: skip
    """"B-prod-like" booster: 27 features, `n_trees` trees, depth <= `max_depth`,
    grown on synthetic fields of a C<grid_n> grid against a synthetic log10(OH) target."""
    raw = raw_fields(grid_n, seed=seed)
    x = quick_features(raw)
    rng = np.random.default_rng(seed + 1)
    if x.shape[0] > n_sample:
        x = x[rng.choice(x.shape[0], n_sample, replace=False)]
    y = synthetic_log10_oh(x, rng)
    return grow_forest(x, y, n_trees=n_trees, max_depth=max_depth, min_leaf=min_leaf, n_cand=n_cand,
                       seed=seed + 2)  # fmt: skip
