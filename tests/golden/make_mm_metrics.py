#!/usr/bin/env python
"""Golden generator for MMADE / MMFDE: runs the reference's own mmade / mmfde (src/metrics/multimodal.py:105-135) on seeded
predictions with ragged multimodal ground-truth groups and stores inputs + results in tests/golden/mm_metrics.npz."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SKELDIFF_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
from src.metrics.multimodal import mmade, mmfde  # noqa: E402

g = torch.Generator().manual_seed(31)
W, S, T, J = 6, 10, 12, 21
pred = torch.rand(W, S, T, J, 3, generator=g) * 2 - 1
counts = [1, 4, 2, 9, 1, 3]
mm_gt = [torch.rand(c, T, J, 3, generator=g) * 2 - 1 for c in counts]
target = torch.rand(W, T, J, 3, generator=g)
a, f = mmade(target, pred, mm_gt), mmfde(target, pred, mm_gt)
path = os.path.join(HERE, "mm_metrics.npz")
np.savez_compressed(path, pred=pred.numpy(), mm_gt=torch.cat(mm_gt, 0).numpy(), counts=np.array(counts), mmade=a.numpy(), mmfde=f.numpy())
print("wrote", path, os.path.getsize(path) // 1024, "KiB", a.tolist())
