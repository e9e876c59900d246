"""Skeleton tables (input provider for the sampling path).

The reference derives the node adjacency (``Kinematic.adj_matrix``,
/root/reference/src/data/skeleton/kinematic/base.py:72-74) and the node-type vector
(``nodes_type_id``, base.py:58-70) from per-dataset joint dictionaries
(kinematic/amass.py:34-70, kinematic/h36m.py:68-111, kinematic/freeman.py:5-43) with
``if_consider_hip=False``.  Only the resulting *data* is needed by the sampling path, so it is
tabulated here (node limb list + type id per node); tests/golden/make_golden.py checks these
tables against the reference classes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

__all__ = ["SkeletonSpec", "get_skeleton", "SKELETONS"]


@dataclass(frozen=True)
class SkeletonSpec:
    name: str
    num_nodes: int
    node_limbseq: Tuple[Tuple[int, int], ...]
    node_types: Tuple[int, ...]
    obs_length: int
    pred_length: int
    pose_box_size: float
    enc_num_layers: int
    node_names: Tuple[str, ...] = ()
    extra: dict = field(default_factory=dict)

    @property
    def nodes_type_id(self) -> torch.Tensor:
        return torch.tensor(self.node_types, dtype=torch.long)

    @property
    def adj_matrix(self) -> torch.Tensor:
        """Symmetric 0/1 adjacency, float32 [N, N] (kinematic/utils.py:4-13)."""
        adj = torch.zeros(self.num_nodes, self.num_nodes)
        for i, j in self.node_limbseq:
            adj[i, j] = 1.0
            adj[j, i] = 1.0
        return adj

    def reachability_matrix(self, factor: float = 0.5, stop_at: Optional[str] = "hips") -> torch.Tensor:
        """`Kinematic.reachability_matrix` (kinematic/base.py:85-127): entry (i, j) = factor ** (d - 1) where d is the length
        found by the reference's depth-first search from i to j, 0 when the search reports j unreachable.  The search is
        reproduced with its quirks, because the covariance of `covariance_matrix_type='reachability'` is defined by it:
        the start node is not in the visited list, neighbours are tried in index order, and meeting a stop node (hips / BMN)
        among the neighbours ends the search of that node with 'unreachable' even if an earlier neighbour had found a path.
        stop_at: 'hips', 'bmn' or None (any other value asserts, like the reference's default stop_at=0)."""
        adj = self.adj_matrix
        N = self.num_nodes
        if stop_at is None:
            stops = None
        elif stop_at == "hips":
            stops = {k for k, v in enumerate(self.node_names) if "hip" in v.lower()}
        elif stop_at == "bmn":
            stops = {k for k, v in enumerate(self.node_names) if "bmn" in v.lower()}
        else:
            raise AssertionError("Not implemented")
        nbrs = [[k for k in range(N) if adj[i, k] == 1] for i in range(N)]

        def search(i: int, j: int, visited: Tuple[int, ...]) -> int:
            if adj[i, j] == 1:
                return 1
            best = 0                                      # 0 = no path found yet
            for k in nbrs[i]:
                if stops is not None and k in stops:
                    return 0
                if k not in visited:
                    d = search(k, j, visited + (k,))
                    if d > 0:
                        best = d + 1 if best == 0 else min(best, d + 1)
            return best

        reach = torch.zeros_like(adj)
        for i in range(N):
            for j in range(i + 1, N):
                d = search(i, j, ())
                reach[i, j] = reach[j, i] = factor ** (d - 1) if d > 0 else 0.0
        return reach

    def transform_to_metric_space(self, kpts: torch.Tensor) -> torch.Tensor:
        """Root-relative unit-box poses -> metres (motion/rescalepose.py:29-39)."""
        return kpts * self.pose_box_size


_AMASS_LIMBS = ((0, 1), (0, 2), (1, 2), (2, 5), (5, 8), (8, 11), (11, 14), (8, 13), (13, 16),
                (16, 18), (18, 20), (8, 12), (12, 15), (15, 17), (17, 19), (1, 4), (4, 7), (7, 10),
                (0, 3), (3, 6), (6, 9))
_H36M_LIMBS = ((0, 3), (0, 6), (3, 6), (0, 1), (1, 2), (3, 4), (4, 5), (6, 7), (7, 8), (8, 9),
               (7, 10), (7, 13), (10, 11), (11, 12), (13, 14), (14, 15))
_FREEMAN_LIMBS = ((1, 0), (1, 6), (0, 6), (0, 2), (1, 3), (2, 4), (3, 5), (6, 7), (6, 8), (7, 9),
                  (8, 10), (6, 11), (6, 12), (11, 13), (12, 14), (13, 15), (14, 16))

_AMASS_NAMES = ("LHip", "RHip", "Spine1", "LKnee", "RKnee", "Spine3", "LHeel", "RHeel", "Neck", "LFoot", "RFoot", "BMN", "LSI", "RSI",
                "Head", "LShoulder", "RShoulder", "LElbow", "RElbow", "LWrist", "RWrist")
_H36M_NAMES = ("RHip", "RKnee", "RAnkle", "LHip", "LKnee", "LAnkle", "Torso", "Neck", "Nose", "Head", "LShoulder", "LElbow", "LWrist",
               "RShoulder", "RElbow", "RWrist")
_FREEMAN_NAMES = ("LHip", "RHip", "LKnee", "RKnee", "LAnkle", "RAnkle", "Nose", "LEye", "REye", "LEar", "REar", "LShoulder", "RShoulder",
                  "LElbow", "RElbow", "LWrist", "RWrist")
# AMASS with MANO hands (kinematic/amass.py:7-85, 52 joints -> 51 nodes): 15 finger joints per hand hang off the wrists; finger
# node names do not follow the L/R + capital convention, so every finger joint is its own node type (13 + 30 = 43 types)
_FINGERS = ("index", "middle", "pinky", "ring", "thumb")
_MANO_NAMES = _AMASS_NAMES + tuple(f"{side}_{f}{i}" for side in ("left", "right") for f in _FINGERS for i in (1, 2, 3))
_MANO_LIMBS = _AMASS_LIMBS + tuple(
    limb for wrist, base in ((19, 21), (20, 36))
    for limb in (tuple((wrist, base + 3 * f) for f in range(5)) + tuple((base + 3 * f + i, base + 3 * f + i + 1) for f in range(5) for i in range(2))))
_MANO_TYPES = (0, 0, 1, 2, 2, 3, 4, 4, 5, 6, 6, 7, 8, 8, 9, 10, 10, 11, 11, 12, 12) + tuple(range(13, 43))

SKELETONS = {
    # configs/config_eval/dataset/amass.yaml, task/hmp.yaml: 30 obs / 120 pred frames; box 1.5 m (config_train_autoencoder/task/hmp.yaml:9)
    "amass": SkeletonSpec("amass", 21, _AMASS_LIMBS,
                          (0, 0, 1, 2, 2, 3, 4, 4, 5, 6, 6, 7, 8, 8, 9, 10, 10, 11, 11, 12, 12),
                          obs_length=30, pred_length=120, pose_box_size=1.5, enc_num_layers=2, node_names=_AMASS_NAMES),
    # configs/config_eval/dataset/amass-mano.yaml: 52 joints incl. hip, fps 60 -> 30 obs / 120 pred
    "amass-mano": SkeletonSpec("amass-mano", 51, _MANO_LIMBS, _MANO_TYPES,
                               obs_length=30, pred_length=120, pose_box_size=1.5, enc_num_layers=2, node_names=_MANO_NAMES),
    # configs/config_eval/dataset/h36m.yaml: fps 50 -> 25 obs / 100 pred
    "h36m": SkeletonSpec("h36m", 16, _H36M_LIMBS,
                         (0, 1, 2, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 7, 8, 9),
                         obs_length=25, pred_length=100, pose_box_size=1.5, enc_num_layers=1, node_names=_H36M_NAMES),
    # configs/config_eval/dataset/freeman.yaml: fps 30 -> 15 obs / 60 pred
    "freeman": SkeletonSpec("freeman", 17, _FREEMAN_LIMBS,
                            (0, 0, 1, 1, 2, 2, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8),
                            obs_length=15, pred_length=60, pose_box_size=1.5, enc_num_layers=1, node_names=_FREEMAN_NAMES),
}


def get_skeleton(name: str) -> SkeletonSpec:
    key = name.lower()
    if key not in SKELETONS:
        raise KeyError(f"unknown skeleton '{name}' (have {sorted(SKELETONS)})")
    return SKELETONS[key]
