#!/usr/bin/env python
"""Golden generator for the long-term (autoregressive) evaluation: runs the reference's own
long_term_prediction_best_every50 (src/eval_utils.py:44-67) with process_evaluation_pair (src/eval_prepare_model.py:124-134) and
get_best_sample_idx (src/metrics/utils.py:22-30), unmodified, around a deterministic stand-in predictor
(skeletondiffusion_b200.testing.fake_predictor), and stores inputs + outputs in tests/golden/long_term.npz.
The reference module never imports `math` although the function uses it; the generator puts it into the module namespace."""
import math
import os
import sys
import warnings
from functools import partial

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SKELDIFF_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF, ROOT]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import src.eval_utils as eu  # noqa: E402
from src.data.skeleton import create_skeleton  # noqa: E402
from src.eval_prepare_model import process_evaluation_pair  # noqa: E402
from skeletondiffusion_b200.testing import fake_predictor  # noqa: E402

eu.math = math
sk = create_skeleton(dataset_name="h36m", motion_repr_type="SkeletonRescalePose", num_joints=17, if_consider_hip=False, obs_length=25,
                     pred_length=100, pose_box_size=1.5)
W, S, n_past, T, J, factor = 5, 7, 6, 10, 16, 2.5
g = torch.Generator().manual_seed(9)
data = (torch.randn(W, n_past, J, 3, generator=g) * 0.3).clamp(-1, 1)
target = (torch.randn(W, int(T * factor), J, 3, generator=g) * 0.3).clamp(-1, 1)      # the reference asserts len(target) == int(factor * T)
predict = fake_predictor(77, S, T)
tgt, pred, mm_gt, obs = eu.long_term_prediction_best_every50(
    data, target, None, get_prediction=lambda d, extra=None: predict(d), process_evaluation_pair=partial(process_evaluation_pair, skeleton=sk),
    num_samples=S, config={"long_term_factor": factor, "pred_length": T})
path = os.path.join(HERE, "long_term.npz")
np.savez_compressed(path, data=data.numpy(), target=target.numpy(), out_target=tgt.numpy(), out_pred=pred.numpy(), out_obs=obs.numpy(),
                    S=S, T=T, factor=factor, seed=77)
print("wrote", path, tuple(pred.shape), tuple(tgt.shape))
