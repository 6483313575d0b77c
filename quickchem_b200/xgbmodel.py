"""In-memory XGBoost forest + model-file *writers* (legacy binary "binf", JSON, UBJSON).

This module is host-side tooling: it builds booster files that the CUDA library
(`libqcoh.so`, `XGBoosterLoadModel`) and the CPU oracle both read.  It never
predicts anything.  The on-disk layouts follow XGBoost 1.6.0, the version the
reference pins (`Shared/CMakeLists.txt:8`, `find_package(xgboost 1.6.0 EXACT)`),
and the file kinds the reference configures (`OH_GridComp/OH_instance_OH.rc:17-20`:
`*.bin` from 0.81 and `*.model` from 1.6.0 — both legacy binary).

Layout notes (little-endian), restated from the published XGBoost 1.6.0 format:
  [ "binf" ] LearnerModelParamLegacy(136 B) str(objective) str(booster)
  GBTreeModelParam(160 B) { TreeParam(148 B) Node[n](20 B) RTreeNodeStat[n](16 B) } x T
  int32 tree_info[T] [attrs] [metrics]
with str = uint64 length + bytes.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np

NODE_DTYPE = np.dtype(
    [("parent", "<i4"), ("cleft", "<i4"), ("cright", "<i4"), ("sindex", "<u4"), ("info", "<f4")]
)
STAT_DTYPE = np.dtype(
    [("loss_chg", "<f4"), ("sum_hess", "<f4"), ("base_weight", "<f4"), ("leaf_child_cnt", "<i4")]
)
assert NODE_DTYPE.itemsize == 20 and STAT_DTYPE.itemsize == 16

JSON_ROOT_PARENT = 2147483647


@dataclass
class Tree:
    """One regression tree in XGBoost's node-array form (node 0 is the root)."""

    left: np.ndarray  # int32, -1 at leaves
    right: np.ndarray  # int32, -1 at leaves
    parent: np.ndarray  # int32, -1 at root (no left-child flag)
    split_index: np.ndarray  # uint32 feature id (0 at leaves)
    split_cond: np.ndarray  # float32 threshold, or leaf value at leaves
    default_left: np.ndarray  # uint8
    sum_hess: np.ndarray | None = None  # float32 cover (optional)
    loss_chg: np.ndarray | None = None
    base_weight: np.ndarray | None = None

    @property
    def num_nodes(self) -> int:
        return int(self.left.shape[0])

    def is_leaf(self) -> np.ndarray:
        return self.left == -1

    def max_depth(self) -> int:
        depth = np.zeros(self.num_nodes, np.int32)
        for n in range(1, self.num_nodes):  # parents precede children in every writer here
            depth[n] = depth[self.parent[n]] + 1
        return int(depth.max())

    def stats(self):
        n = self.num_nodes
        sh = self.sum_hess if self.sum_hess is not None else np.zeros(n, np.float32)
        lc = self.loss_chg if self.loss_chg is not None else np.zeros(n, np.float32)
        bw = self.base_weight if self.base_weight is not None else np.zeros(n, np.float32)
        return sh.astype(np.float32), lc.astype(np.float32), bw.astype(np.float32)


@dataclass
class Forest:
    trees: List[Tree] = field(default_factory=list)
    base_score: float = 0.5
    num_feature: int = 27
    objective: str = "reg:squarederror"
    version: tuple = (1, 6, 0)
    attributes: dict = field(default_factory=dict)

    @property
    def num_trees(self) -> int:
        return len(self.trees)

    def total_nodes(self) -> int:
        return sum(t.num_nodes for t in self.trees)


# --------------------------------------------------------------------------------------
# small constructors used by tests (hand-derivable known-answer boosters)
# --------------------------------------------------------------------------------------
def tree_from_nested(spec) -> Tree:
    """Build a Tree from a nested spec.

    spec := float                                  (leaf value)
          | (feature, threshold, default_left, left_spec, right_spec)
    Nodes are numbered the way XGBoost grows them: children are allocated as an adjacent
    pair (cright == cleft + 1) when their parent is expanded, breadth-first.
    """
    left, right, parent, sidx, cond, dl = [], [], [], [], [], []

    def new_node(p):
        left.append(-1), right.append(-1), parent.append(p), sidx.append(0), cond.append(0.0), dl.append(0)
        return len(left) - 1

    from collections import deque

    queue = deque([(new_node(-1), spec)])
    while queue:
        nid, sp = queue.popleft()
        if isinstance(sp, (int, float, np.floating)):
            cond[nid] = float(sp)
            continue
        f, thr, d, ls, rs = sp
        l = new_node(nid)
        r = new_node(nid)
        left[nid], right[nid], sidx[nid], cond[nid], dl[nid] = l, r, int(f), float(thr), int(bool(d))
        queue.append((l, ls))
        queue.append((r, rs))
    return Tree(
        left=np.asarray(left, np.int32),
        right=np.asarray(right, np.int32),
        parent=np.asarray(parent, np.int32),
        split_index=np.asarray(sidx, np.uint32),
        split_cond=np.asarray(cond, np.float32),
        default_left=np.asarray(dl, np.uint8),
    )


# --------------------------------------------------------------------------------------
# legacy binary writer
# --------------------------------------------------------------------------------------
def _wstr(b: bytearray, s: str) -> None:
    raw = s.encode()
    b += struct.pack("<Q", len(raw)) + raw


def legacy_binary_bytes(forest: Forest, with_binf: bool = True, objective: str | None = None) -> bytes:
    """Serialise `forest` the way XGBoost 1.6.0 `LearnerIO::SaveModel(dmlc::Stream*)` does."""
    b = bytearray()
    if with_binf:
        b += b"binf"
    contain_attrs = 1 if forest.attributes else 0
    major, minor = forest.version[0], forest.version[1]
    # LearnerModelParamLegacy: base_score, num_feature, num_class, contain_extra_attrs,
    # contain_eval_metrics, major, minor, num_target, reserved[26]  -> 136 B
    b += struct.pack("<fIiiiIII", forest.base_score, forest.num_feature, 0, contain_attrs, 0, major, minor, 1)
    b += b"\0" * (26 * 4)
    _wstr(b, objective or forest.objective)
    _wstr(b, "gbtree")
    # GBTreeModelParam: num_trees, num_roots, num_feature, pad, int64 num_pbuffer,
    # num_output_group, size_leaf_vector, reserved[32] -> 160 B
    b += struct.pack("<iiiiqii", forest.num_trees, 1, forest.num_feature, 0, 0, 1, 0)
    b += b"\0" * (32 * 4)
    for t in forest.trees:
        n = t.num_nodes
        # TreeParam: num_roots, num_nodes, num_deleted, max_depth, num_feature, size_leaf_vector, reserved[31]
        b += struct.pack("<iiiiii", 1, n, 0, 0, forest.num_feature, 0)
        b += b"\0" * (31 * 4)
        nodes = np.zeros(n, NODE_DTYPE)
        par = t.parent.astype(np.int64).copy()
        # high bit of parent_ marks "is left child" (RegTree::Node::SetParent)
        is_left = np.zeros(n, bool)
        internal = np.nonzero(t.left != -1)[0]
        is_left[t.left[internal]] = True
        enc = np.where(par < 0, -1, np.where(is_left, par | (1 << 31), par))
        nodes["parent"] = (enc & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
        nodes["cleft"] = t.left
        nodes["cright"] = t.right
        sidx = t.split_index.astype(np.uint32) | (t.default_left.astype(np.uint32) << 31)
        sidx = np.where(t.left == -1, np.uint32(0), sidx).astype(np.uint32)
        nodes["sindex"] = sidx
        nodes["info"] = t.split_cond.astype(np.float32)
        b += nodes.tobytes()
        st = np.zeros(n, STAT_DTYPE)
        sh, lc, bw = t.stats()
        st["sum_hess"], st["loss_chg"], st["base_weight"] = sh, lc, bw
        b += st.tobytes()
    b += np.zeros(forest.num_trees, "<i4").tobytes()  # tree_info: output group 0
    if contain_attrs:
        items = sorted(forest.attributes.items())
        b += struct.pack("<Q", len(items))
        for k, v in items:
            _wstr(b, k)
            _wstr(b, v)
    return bytes(b)


def write_legacy_binary(forest: Forest, path: str, with_binf: bool = True, objective: str | None = None) -> None:
    with open(path, "wb") as f:
        f.write(legacy_binary_bytes(forest, with_binf, objective))


def read_legacy_binary(path: str) -> Forest:
    """Inverse of `write_legacy_binary` for files this module wrote (tooling: the booster sweep replicates grown
    forests; the library and the oracle have their own, stricter readers)."""
    raw = open(path, "rb").read()
    o = 4 if raw[:4] == b"binf" else 0
    base_score, num_feature = struct.unpack_from("<fI", raw, o)
    o += 136

    def rstr():
        nonlocal o
        (n,) = struct.unpack_from("<Q", raw, o)
        o += 8 + n
        return raw[o - n : o].decode()

    objective, booster = rstr(), rstr()
    assert booster == "gbtree", booster
    (num_trees,) = struct.unpack_from("<i", raw, o)
    o += 160
    trees = []
    for _ in range(num_trees):
        (n,) = struct.unpack_from("<i", raw, o + 4)
        o += 148
        nodes = np.frombuffer(raw, NODE_DTYPE, n, o)
        o += 20 * n
        st = np.frombuffer(raw, STAT_DTYPE, n, o)
        o += 16 * n
        par = nodes["parent"].astype(np.int64)
        par = np.where(par == -1, -1, par & 0x7FFFFFFF)
        trees.append(Tree(left=nodes["cleft"].astype(np.int32), right=nodes["cright"].astype(np.int32),
                          parent=par.astype(np.int32), split_index=(nodes["sindex"] & 0x7FFFFFFF).astype(np.uint32),
                          split_cond=nodes["info"].astype(np.float32), default_left=(nodes["sindex"] >> 31).astype(np.uint8),
                          sum_hess=st["sum_hess"].astype(np.float32)))  # fmt: skip
    return Forest(trees=trees, base_score=float(np.float32(base_score)), num_feature=int(num_feature), objective=objective)


# --------------------------------------------------------------------------------------
# JSON / UBJSON writers (XGBoost 1.6 schema, doc/model.schema)
# --------------------------------------------------------------------------------------
def _f32_repr(x: np.float32) -> float:
    """A Python float that round-trips through float32 exactly and prints shortest."""
    return float(np.format_float_scientific(np.float32(x), unique=True))


def forest_to_jsonable(forest: Forest) -> dict:
    trees = []
    for i, t in enumerate(forest.trees):
        n = t.num_nodes
        sh, lc, bw = t.stats()
        par = np.where(t.parent < 0, JSON_ROOT_PARENT, t.parent).astype(np.int64)
        trees.append(
            {
                "base_weights": [_f32_repr(v) for v in bw],
                "categories": [],
                "categories_nodes": [],
                "categories_segments": [],
                "categories_sizes": [],
                "default_left": [int(v) for v in t.default_left],
                "id": i,
                "left_children": [int(v) for v in t.left],
                "loss_changes": [_f32_repr(v) for v in lc],
                "parents": [int(v) for v in par],
                "right_children": [int(v) for v in t.right],
                "split_conditions": [_f32_repr(v) for v in t.split_cond],
                "split_indices": [int(v) for v in t.split_index],
                "split_type": [0] * n,
                "sum_hessian": [_f32_repr(v) for v in sh],
                "tree_param": {
                    "num_deleted": "0",
                    "num_feature": str(forest.num_feature),
                    "num_nodes": str(n),
                    "size_leaf_vector": "0",
                },
            }
        )
    return {
        "learner": {
            "attributes": dict(forest.attributes),
            "feature_names": [],
            "feature_types": [],
            "gradient_booster": {
                "model": {
                    "gbtree_model_param": {
                        "num_parallel_tree": "1",
                        "num_trees": str(forest.num_trees),
                        "size_leaf_vector": "0",
                    },
                    "tree_info": [0] * forest.num_trees,
                    "trees": trees,
                },
                "name": "gbtree",
            },
            "learner_model_param": {
                "base_score": np.format_float_scientific(np.float32(forest.base_score), unique=True).upper(),
                "num_class": "0",
                "num_feature": str(forest.num_feature),
                "num_target": "1",
            },
            "objective": {"name": forest.objective, "reg_loss_param": {"scale_pos_weight": "1"}},
        },
        "version": list(forest.version),
    }


def write_json(forest: Forest, path: str) -> None:
    with open(path, "w") as f:
        json.dump(forest_to_jsonable(forest), f, separators=(",", ":"))


def _ubj_str(s: str) -> bytes:
    raw = s.encode()
    return b"L" + struct.pack(">q", len(raw)) + raw


def _ubj_typed(arr: np.ndarray, marker: bytes, fmt: str) -> bytes:
    return b"[$" + marker + b"#L" + struct.pack(">q", arr.size) + arr.astype(fmt).tobytes()


def _ubj(obj) -> bytes:
    """Universal Binary JSON (draft 12) as XGBoost 1.6 `UBJWriter` emits it (big-endian)."""
    if isinstance(obj, dict):
        out = b"{"
        for k, v in obj.items():
            out += _ubj_str(k) + _ubj(v)
        return out + b"}"
    if isinstance(obj, np.ndarray):
        if obj.dtype == np.float32:
            return _ubj_typed(obj, b"d", ">f4")
        if obj.dtype == np.uint8:
            return _ubj_typed(obj, b"U", ">u1")
        if obj.dtype == np.int32:
            return _ubj_typed(obj, b"l", ">i4")
        if obj.dtype == np.int64:
            return _ubj_typed(obj, b"L", ">i8")
        raise TypeError(obj.dtype)
    if isinstance(obj, (list, tuple)):
        return b"[" + b"".join(_ubj(v) for v in obj) + b"]"
    if isinstance(obj, str):
        return b"S" + _ubj_str(obj)
    if isinstance(obj, bool):
        return b"T" if obj else b"F"
    if isinstance(obj, int):
        return b"L" + struct.pack(">q", obj)
    if isinstance(obj, float):
        return b"d" + struct.pack(">f", obj)
    raise TypeError(type(obj))


def write_ubj(forest: Forest, path: str) -> None:
    j = forest_to_jsonable(forest)
    model = j["learner"]["gradient_booster"]["model"]
    model["tree_info"] = np.asarray(model["tree_info"], np.int32)
    for tj, t in zip(model["trees"], forest.trees):
        sh, lc, bw = t.stats()
        tj["base_weights"] = bw
        tj["loss_changes"] = lc
        tj["sum_hessian"] = sh
        tj["split_conditions"] = t.split_cond.astype(np.float32)
        tj["default_left"] = t.default_left.astype(np.uint8)
        tj["left_children"] = t.left.astype(np.int32)
        tj["right_children"] = t.right.astype(np.int32)
        tj["parents"] = np.where(t.parent < 0, JSON_ROOT_PARENT, t.parent).astype(np.int32)
        tj["split_indices"] = t.split_index.astype(np.int32)
        tj["split_type"] = np.zeros(t.num_nodes, np.uint8)
        for k in ("categories", "categories_nodes", "categories_segments", "categories_sizes"):
            tj[k] = np.zeros(0, np.int32)
    with open(path, "wb") as f:
        f.write(_ubj(j))
