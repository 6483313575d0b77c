"""Deliberately naive second implementation of the XGBoost 1.6.0 predict rules (numpy / pure
Python), plus independent JSON / UBJSON / legacy-binary model *readers*.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/qc_oracle.h).  It exists so that three
independently written traversers (this one, oracle/qc_oracle.c, and the CUDA kernels) must
agree bit-for-bit on leaf ids.

Rules restated (xgboost 1.6.0 src/predictor/predict_fn.h `GetNextNode`,
src/predictor/cpu_predictor.cc `PredictByAllTrees`, src/data/data.cc `SparsePage::Push`):
  * an input entry is missing iff isnan(v) or v == missing (the reference passes -999.0,
    OH_GridComp/OH_GridCompMod.F90:213,347)
  * at a split: missing -> default child; else left + !(fvalue < split_cond) in float32
  * out = base_score; out += leaf_value tree by tree, in float32
"""
from __future__ import annotations

import json
import struct

import numpy as np

ROOT_PARENT = 2147483647


class NaiveTree:
    def __init__(self, left, right, split_index, split_cond, default_left):
        self.left = np.asarray(left, np.int64)
        self.right = np.asarray(right, np.int64)
        self.split_index = np.asarray(split_index, np.int64)
        self.split_cond = np.asarray(split_cond, np.float32)
        self.default_left = np.asarray(default_left, np.int64)


class NaiveModel:
    def __init__(self, trees, base_score, num_feature, objective):
        self.trees, self.base_score, self.num_feature, self.objective = trees, np.float32(base_score), num_feature, objective


def read_json(path) -> NaiveModel:
    with open(path) as f:
        j = json.load(f)
    return _from_jsonable(j)


def _from_jsonable(j) -> NaiveModel:
    lr = j["learner"]
    trees = []
    for t in lr["gradient_booster"]["model"]["trees"]:
        trees.append(NaiveTree(t["left_children"], t["right_children"], t["split_indices"],
                               np.asarray(t["split_conditions"], np.float64).astype(np.float32),
                               [int(v) for v in t["default_left"]]))  # fmt: skip
    p = lr["learner_model_param"]
    return NaiveModel(trees, np.float32(float(p["base_score"])), int(p["num_feature"]), lr["objective"]["name"])


def read_ubj(path) -> NaiveModel:
    data = open(path, "rb").read()
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, data, pos)
        pos += struct.calcsize(fmt)
        return v[0]

    scal = {b"i": ">b", b"U": ">B", b"I": ">h", b"l": ">i", b"L": ">q", b"d": ">f", b"D": ">d"}
    npdt = {b"i": ">i1", b"U": ">u1", b"I": ">i2", b"l": ">i4", b"L": ">i8", b"d": ">f4", b"D": ">f8"}

    def rint():
        m = data[pos : pos + 1]
        return rd_marker_value(m, advance=True)

    def rd_marker_value(m, advance):
        nonlocal pos
        if advance:
            pos += 1
        return rd(scal[m])

    def rstr():
        n = rint()
        nonlocal pos
        s = data[pos : pos + n].decode()
        pos += n
        return s

    def value(m=None):
        nonlocal pos
        if m is None:
            m = data[pos : pos + 1]
            pos += 1
        if m == b"{":
            out = {}
            while data[pos : pos + 1] != b"}":
                k = rstr()
                out[k] = value()
            pos += 1
            return out
        if m == b"[":
            if data[pos : pos + 1] == b"$":
                ty = data[pos + 1 : pos + 2]
                assert data[pos + 2 : pos + 3] == b"#"
                pos += 3
                n = rint()
                a = np.frombuffer(data, npdt[ty], n, pos)
                pos += a.nbytes
                return a.astype(a.dtype.newbyteorder("="))
            out = []
            while data[pos : pos + 1] != b"]":
                out.append(value())
            pos += 1
            return out
        if m == b"S":
            return rstr()
        if m == b"T":
            return True
        if m == b"F":
            return False
        if m == b"Z":
            return None
        return rd_marker_value(m, advance=False)

    return _from_jsonable(value())


def read_legacy(path) -> NaiveModel:
    buf = open(path, "rb").read()
    pos = 4 if buf[:4] == b"binf" else 0
    base_score, num_feature = struct.unpack_from("<fI", buf, pos)
    pos += 136

    def rstr():
        nonlocal pos
        (n,) = struct.unpack_from("<Q", buf, pos)
        s = buf[pos + 8 : pos + 8 + n].decode()
        pos += 8 + n
        return s

    objective, booster = rstr(), rstr()
    assert booster == "gbtree"
    (num_trees,) = struct.unpack_from("<i", buf, pos)
    pos += 160
    node_dt = np.dtype([("parent", "<i4"), ("cleft", "<i4"), ("cright", "<i4"), ("sindex", "<u4"), ("info", "<f4")])
    trees = []
    for _ in range(num_trees):
        (num_nodes,) = struct.unpack_from("<i", buf, pos + 4)
        pos += 148
        nd = np.frombuffer(buf, node_dt, num_nodes, pos)
        pos += 36 * num_nodes
        trees.append(NaiveTree(nd["cleft"], nd["cright"], nd["sindex"] & 0x7FFFFFFF, nd["info"], nd["sindex"] >> 31))
    return NaiveModel(trees, base_score, num_feature, objective)


def leaf_ids(model: NaiveModel, x: np.ndarray, missing=np.float32(-999.0), ntree_limit=0) -> np.ndarray:
    """[nrow][ntree] int64 leaf node ids, one Python loop iteration per tree level (vectorised
    over rows).  Absent columns (ncol < num_feature) count as missing."""
    x = np.asarray(x, np.float32)
    nrow, ncol = x.shape
    miss = np.isnan(x) | (x == np.float32(missing))
    nt = len(model.trees) if ntree_limit == 0 else min(ntree_limit, len(model.trees))
    out = np.zeros((nrow, nt), np.int64)
    rows = np.arange(nrow)
    for ti in range(nt):
        t = model.trees[ti]
        nid = np.zeros(nrow, np.int64)
        while True:
            active = t.left[nid] != -1
            if not active.any():
                break
            f = t.split_index[nid]
            fin = np.minimum(f, ncol - 1)
            is_miss = miss[rows, fin] | (f >= ncol)
            fv = x[rows, fin]
            nxt = np.where(is_miss, np.where(t.default_left[nid] != 0, t.left[nid], t.right[nid]),
                           t.left[nid] + (~(fv < t.split_cond[nid])).astype(np.int64))  # fmt: skip
            nid = np.where(active, nxt, nid)
        out[:, ti] = nid
    return out


def predict(model: NaiveModel, x, missing=np.float32(-999.0), ntree_limit=0) -> np.ndarray:
    ids = leaf_ids(model, x, missing, ntree_limit)
    out = np.full(ids.shape[0], model.base_score, np.float32)
    for ti in range(ids.shape[1]):
        out = (out + model.trees[ti].split_cond[ids[:, ti]]).astype(np.float32)
    return out


def predict_scalar(model: NaiveModel, row, missing=-999.0):
    """Pure-Python single-row walk (for hand-checkable cases)."""
    out = np.float32(model.base_score)
    ids = []
    for t in model.trees:
        nid = 0
        while t.left[nid] != -1:
            f = int(t.split_index[nid])
            v = np.float32(row[f]) if f < len(row) else np.float32(np.nan)
            if np.isnan(v) or v == np.float32(missing):
                nid = int(t.left[nid] if t.default_left[nid] else t.right[nid])
            else:
                nid = int(t.left[nid]) + (0 if v < t.split_cond[nid] else 1)
        ids.append(nid)
        out = np.float32(out + t.split_cond[nid])
    return out, ids
