"""ADE / FDE / APD of a batch of predicted motions on the GPU (one kernel launch, predictions read once).

Mirrors `ade`, `fde`, `apd` of the reference (src/metrics/multimodal.py:44-57, :60-73, :15-35): same argument order
(`target, pred` / `pred`), same `t0` / `t` frame window (multimodal.py:4-8), one value per observed window.  `eval.py`
calls them on `skeleton.transform_to_metric_space(...)` (rescalepose.py:29-39), a multiplication by `pose_box_size`
that all three metrics are linear in: pass it as `scale` to `motion_metrics` instead of scaling the 1.5 MB / window
prediction tensor.  CUDA only: there is no CPU fallback (the CPU statement of the metrics lives in the test infrastructure)."""
from __future__ import annotations

import itertools
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _native as nv


def _frames(x: torch.Tensor, t0: int, t: int, axis: int) -> torch.Tensor:
    end = x.shape[axis] if t == -1 else t               # multimodal.py:4-8
    return x if (t0 == 0 and end == x.shape[axis]) else x.narrow(axis, t0, end - t0)


def motion_metrics(target: Optional[torch.Tensor], pred: torch.Tensor, scale: float = 1.0, t0: int = 0, t: int = -1,
                   want: Tuple[bool, bool, bool] = (True, True, True)):
    """pred [W, S, T, ...], target [W, T, ...] (fp32, CUDA) -> (ade [W], fde [W], apd [W]); entries not in `want` are None.
    `target` may be None when only APD is wanted."""
    nv.require_cuda(pred, "pred")
    if pred.dim() < 4:
        raise ValueError(f"pred must be [windows, samples, frames, ...], got {tuple(pred.shape)}")
    want_ade, want_fde, want_apd = want
    pred = _frames(pred, t0, t, 2)
    W, S, T = pred.shape[:3]
    F = math.prod(pred.shape[3:])
    p = pred.reshape(W, S, T, F).to(torch.float32).contiguous()
    if target is None:
        if want_ade or want_fde:
            raise ValueError("ADE / FDE need a target")
        tg = p[:, 0]                                   # read but unused by the APD result
    else:
        nv.require_cuda(target, "target")
        tg = _frames(target, t0, t, 1)
        if tg.shape[0] != W or tg.shape[1] != T or math.prod(tg.shape[2:]) != F:
            raise ValueError(f"target {tuple(target.shape)} does not match pred {tuple(pred.shape)}")
        tg = tg.reshape(W, T, F).to(torch.float32).contiguous()
    outs = [torch.empty(W, device=p.device, dtype=torch.float32) if w else None for w in (want_ade, want_fde, want_apd)]
    if W and T and F:
        nv.check(nv.load().sd_motion_metrics(p.data_ptr(), tg.data_ptr(), W, S, T, F, float(scale),
                                             *(nv.dptr(o) for o in outs), nv.stream_ptr(p.device)), "sd_motion_metrics")
    return tuple(outs)


def ade(target, pred, t0=0, t=-1, reduction="mean", **kwargs):
    if reduction != "mean":
        raise NotImplementedError("ade(reduction != 'mean') (per-sample distances) is not on the GPU path")
    return motion_metrics(target, pred, 1.0, t0, t, (True, False, False))[0]


def fde(target, pred, t0=0, t=-1, reduction="mean", **kwargs):
    if reduction != "mean":
        raise NotImplementedError("fde(reduction != 'mean') (per-sample distances) is not on the GPU path")
    return motion_metrics(target, pred, 1.0, t0, t, (False, True, False))[1]


def apd(pred, t0=0, t=-1, **kwargs):
    return motion_metrics(None, pred, 1.0, t0, t, (False, False, True))[2]


def multimodal_metrics(pred: torch.Tensor, mm_gt: Sequence[torch.Tensor], scale: float = 1.0, t0: int = 0, t: int = -1):
    """(mmade [W], mmfde [W]) of pred [W, S, T, ...] against the ragged multimodal ground truths mm_gt[i] [n_i, T, ...]
    (src/metrics/multimodal.py:105-135), one launch of the metric kernel over all ground truths."""
    nv.require_cuda(pred, "pred")
    pred = _frames(pred, t0, t, 2)
    W, S, T = pred.shape[:3]
    F = math.prod(pred.shape[3:])
    p = pred.reshape(W, S, T, F).to(torch.float32).contiguous()
    if len(mm_gt) != W:
        raise ValueError(f"mm_gt must hold one tensor per window ({W}), got {len(mm_gt)}")
    counts = [int(g.shape[0]) for g in mm_gt]
    n_gt = sum(counts)
    dev = p.device
    offsets = torch.tensor([0] + list(itertools.accumulate(counts)), dtype=torch.int32, device=dev)
    out_a, out_f = torch.empty(W, device=dev), torch.empty(W, device=dev)
    if n_gt:
        gts = torch.cat([_frames(g.to(dev, torch.float32), t0, t, 1).reshape(g.shape[0], T, F) for g in mm_gt if g.shape[0]], 0).contiguous()
        win = torch.repeat_interleave(torch.arange(W, dtype=torch.int32, device=dev), torch.tensor(counts, device=dev))
        scratch = torch.empty(2 * n_gt, device=dev)
    else:
        gts = win = scratch = None
    if W:
        nv.check(nv.load().sd_multimodal_metrics(p.data_ptr(), nv.dptr(gts), nv.dptr(win), offsets.data_ptr(), W, n_gt, S, T, F, float(scale),
                                                 out_a.data_ptr(), out_f.data_ptr(), nv.dptr(scratch), nv.stream_ptr(dev)), "sd_multimodal_metrics")
    return out_a, out_f


def mmade(target, pred, mm_gt, t0=0, t=-1, **kwargs):
    return multimodal_metrics(pred, mm_gt, t0=t0, t=t)[0]


def mmfde(target, pred, mm_gt, t0=0, t=-1, **kwargs):
    return multimodal_metrics(pred, mm_gt, t0=t0, t=t)[1]
