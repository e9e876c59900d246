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
from typing import List, Tuple

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

SKELETONS = {
    # configs/config_eval/dataset/amass.yaml, task/hmp.yaml: 30 obs / 120 pred frames; box 1.5 m (config_train_autoencoder/task/hmp.yaml:9)
    "amass": SkeletonSpec("amass", 21, _AMASS_LIMBS,
                          (0, 0, 1, 2, 2, 3, 4, 4, 5, 6, 6, 7, 8, 8, 9, 10, 10, 11, 11, 12, 12),
                          obs_length=30, pred_length=120, pose_box_size=1.5, enc_num_layers=2),
    # configs/config_eval/dataset/h36m.yaml: fps 50 -> 25 obs / 100 pred
    "h36m": SkeletonSpec("h36m", 16, _H36M_LIMBS,
                         (0, 1, 2, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 7, 8, 9),
                         obs_length=25, pred_length=100, pose_box_size=1.5, enc_num_layers=1),
    # configs/config_eval/dataset/freeman.yaml: fps 30 -> 15 obs / 60 pred
    "freeman": SkeletonSpec("freeman", 17, _FREEMAN_LIMBS,
                            (0, 0, 1, 1, 2, 2, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8),
                            obs_length=15, pred_length=60, pose_box_size=1.5, enc_num_layers=1),
}


def get_skeleton(name: str) -> SkeletonSpec:
    key = name.lower()
    if key not in SKELETONS:
        raise KeyError(f"unknown skeleton '{name}' (have {sorted(SKELETONS)})")
    return SKELETONS[key]
