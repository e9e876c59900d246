#!/usr/bin/env python
"""Golden generator for the skeleton tables: runs the UNMODIFIED reference kinematic classes (/root/reference,
src/data/skeleton/kinematic/base.py:58-127) and stores node names / types / adjacency / reachability matrices for the four
skeletons (AMASS, AMASS-MANO, H36M, FreeMan) in tests/golden/skeletons.npz.  Build container only:
    python tests/golden/make_reachability.py"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SKELDIFF_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from src.data.skeleton import create_skeleton  # noqa: E402

out = {}
for name, nj in (("amass", 22), ("amass-mano", 52), ("h36m", 17), ("freeman", 18)):
    sk = create_skeleton(dataset_name=name, motion_repr_type="SkeletonRescalePose", num_joints=nj, if_consider_hip=False,
                         obs_length=30, pred_length=120, pose_box_size=1.2)
    key = name.replace("-", "_")
    out[f"{key}__names"] = np.array(list(sk.node_dict.values()))
    out[f"{key}__types"] = sk.nodes_type_id.numpy()
    out[f"{key}__adj"] = sk.adj_matrix.numpy()
    for stop in ("hips", "bmn", None):
        out[f"{key}__reach_{stop}"] = sk.reachability_matrix(factor=0.5, stop_at=stop).numpy()
    out[f"{key}__reach_hips_f03"] = sk.reachability_matrix(factor=0.3, stop_at="hips").numpy()
    print(name, sk.num_nodes)
path = os.path.join(HERE, "skeletons.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
