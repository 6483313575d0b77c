"""An independent, third-party implementation as an extra anchor for the traversal rules: scikit-learn
grows regression trees and its own `apply()` (Cython, float32 X, `x <= threshold` with float64
thresholds) assigns leaves; the trees are converted to XGBoost's form (`x < thr`, float32) and the
oracle — and, on the GPU, libqcoh — must put every row in the corresponding leaf.

sklearn is NOT libxgboost: this pins "descend by comparing one feature per node, children adjacent,
leaf values summed" against code nobody in this repo wrote; missing values / default directions are
outside sklearn's model and stay pinned by the hand-derived cases only (parity unpinned)."""
import numpy as np
import pytest

from quickchem_b200 import synth, xgbmodel

sk = pytest.importorskip("sklearn.tree")


def sklearn_to_xgb(tree):
    """sklearn Tree -> xgbmodel.Tree (breadth-first, children adjacent) + map xgb node id -> sklearn node id."""
    t = tree.tree_
    left, right, parent, sidx, cond, dl, skid = [-1], [-1], [-1], [0], [0.0], [0], [0]
    queue = [0]
    while queue:
        nid = queue.pop(0)
        s = skid[nid]
        if t.children_left[s] == -1:
            cond[nid] = float(np.float32(t.value[s, 0, 0]))
            continue
        # x <= thr64  <=>  x < nextafter(largest float32 <= thr64, +inf)   for float32 x
        thr64 = t.threshold[s]
        t32 = np.float32(thr64)
        if float(t32) > thr64:
            t32 = np.nextafter(t32, np.float32(-np.inf), dtype=np.float32)
        xthr = np.nextafter(t32, np.float32(np.inf), dtype=np.float32)
        l = len(left)
        for child in (t.children_left[s], t.children_right[s]):
            left.append(-1), right.append(-1), parent.append(nid), sidx.append(0), cond.append(0.0), dl.append(0)
            skid.append(int(child))
        left[nid], right[nid], sidx[nid], cond[nid] = l, l + 1, int(t.feature[s]), float(xthr)
        queue += [l, l + 1]
    xt = xgbmodel.Tree(left=np.asarray(left, np.int32), right=np.asarray(right, np.int32),
                       parent=np.asarray(parent, np.int32), split_index=np.asarray(sidx, np.uint32),
                       split_cond=np.asarray(cond, np.float32), default_left=np.asarray(dl, np.uint8))  # fmt: skip
    return xt, np.asarray(skid)


@pytest.fixture(scope="module")
def sk_forest():
    rng = np.random.default_rng(3)
    x = synth.quick_features(synth.raw_fields(6))
    y = synth.synthetic_log10_oh(x, rng)
    trees, maps, models = [], [], []
    for i in range(5):
        m = sk.DecisionTreeRegressor(max_depth=9, min_samples_leaf=4, random_state=i, splitter="random" if i % 2 else "best")
        m.fit(x[i::5], y[i::5] - (0.0 if i == 0 else 0.1 * i))
        xt, skid = sklearn_to_xgb(m)
        trees.append(xt), maps.append(skid), models.append(m)
    xq = synth.quick_features(synth.raw_fields(8, seed=99))
    # rows exactly on sklearn's thresholds (as float32) — the `<=` vs `<` conversion must hold there
    for m in models:
        t = m.tree_
        internal = np.nonzero(t.children_left != -1)[0][:40]
        for j, s in enumerate(internal):
            xq[(7 * j) % xq.shape[0], t.feature[s]] = np.float32(t.threshold[s])
    return xgbmodel.Forest(trees=trees, base_score=0.0, num_feature=27), maps, models, xq


def _check(leaf, maps, models, xq):
    for i, (skid, m) in enumerate(zip(maps, models)):
        assert np.array_equal(skid[leaf[:, i].astype(np.int64)], m.apply(xq)), f"tree {i}"


def test_oracle_matches_sklearn_apply(oracle, sk_forest, tmp_path):
    forest, maps, models, xq = sk_forest
    p = str(tmp_path / "sk.model")
    xgbmodel.write_legacy_binary(forest, p)
    om = oracle.Model(p)
    _check(om.predict(xq, option_mask=2), maps, models, xq)
    # summed prediction = float32 sum of sklearn's per-tree predictions in tree order
    acc = np.zeros(xq.shape[0], np.float32)
    for m in models:
        acc = (acc + m.predict(xq).astype(np.float32)).astype(np.float32)
    assert np.array_equal(om.predict(xq).view(np.uint32), acc.view(np.uint32))


@pytest.mark.gpu
def test_gpu_matches_sklearn_apply(capi, sk_forest, tmp_path):
    forest, maps, models, xq = sk_forest
    p = str(tmp_path / "sk.model")
    xgbmodel.write_legacy_binary(forest, p)
    b = capi.Booster(p)
    _check(b.predict(capi.DMatrix(xq), option_mask=2), maps, models, xq)
