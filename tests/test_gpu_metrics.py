"""GPU parity of the evaluation-metric kernel (sd_motion_metrics, through the C ABI) against
(i) the ADE / FDE / APD values the reference's own functions (src/metrics/multimodal.py) produced for the golden cases,
(ii) the CPU oracle on seeded inputs, including the reference's edge cases (one sample, one frame, a frame window), and
(iii) size-independent properties at the bench shape (512 windows x 50 samples x 120 frames x 63 features).
Tolerance: 1e-5 relative (fp32 sums of up to 7 560 squared differences; the oracle's torch.cdist takes the
|a|^2 + |b|^2 - 2ab route for more than 25 samples, which is itself ~1e-6 off a float64 evaluation); eval.py prints 4 decimals."""
import pytest
import torch

from oracle import skeldiff_oracle as oc
from tests import _golden as G

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _close(a, b, tol=TOL):
    return float((a.double().cpu() - b.double().cpu()).abs().max()) <= tol * max(1e-30, float(b.double().abs().max()))


@pytest.mark.parametrize("name", ["amass_perturbed", "h36m_perturbed", "freeman_perturbed"])
def test_metrics_match_reference_goldens(cuda_device, name):
    import skeletondiffusion_b200 as sdb
    case = G.load_npz(name)
    spec = sdb.get_skeleton(str(case["dataset"]))
    pred, target = case["pred"].to(cuda_device), case["target"].to(cuda_device)
    a, f, d = sdb.motion_metrics(target, pred, scale=float(spec.pose_box_size))
    assert _close(a, case["ade"]) and _close(f, case["fde"]) and _close(d, case["apd"])
    # reference-named entry points on tensors already in metric space
    pm, tm = spec.transform_to_metric_space(pred), spec.transform_to_metric_space(target)
    assert _close(sdb.ade(tm, pm), case["ade"]) and _close(sdb.fde(tm, pm), case["fde"]) and _close(sdb.apd(pm), case["apd"])


@pytest.mark.parametrize("W,S,T,J", [(3, 50, 120, 21), (5, 1, 7, 16), (4, 2, 1, 17), (2, 91, 5, 21), (7, 13, 33, 5), (1, 50, 100, 64)])
def test_metrics_vs_oracle_seeded(cuda_device, W, S, T, J):
    import skeletondiffusion_b200 as sdb
    g = torch.Generator().manual_seed(W * 1000 + S)
    pred = torch.randn(W, S, T, J, 3, generator=g)
    target = torch.randn(W, T, J, 3, generator=g)
    a, f, d = sdb.motion_metrics(target.to(cuda_device), pred.to(cuda_device))
    assert _close(a, oc.ade(target, pred)) and _close(f, oc.fde(target, pred))
    if S == 1:
        assert float(d.abs().max()) == 0.0                                  # multimodal.py:19-20
    else:
        p64 = pred.double().reshape(W, S, -1)
        iu = torch.triu_indices(S, S, offset=1)
        direct = (p64[:, iu[0]] - p64[:, iu[1]]).norm(dim=-1).mean(-1)
        assert _close(d, direct) and _close(d, oc.apd(pred), 1e-4)
    # bitwise repeatable (fixed-order reductions)
    a2, f2, d2 = sdb.motion_metrics(target.to(cuda_device), pred.to(cuda_device))
    assert torch.equal(a, a2) and torch.equal(f, f2) and torch.equal(d, d2)


def test_metrics_frame_window_and_errors(cuda_device):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200._native import NativeError
    g = torch.Generator().manual_seed(9)
    pred, target = torch.randn(3, 6, 20, 21, 3, generator=g), torch.randn(3, 20, 21, 3, generator=g)
    pc, tc = pred.to(cuda_device), target.to(cuda_device)
    assert _close(sdb.ade(tc, pc, t0=4, t=15), oc.ade(target[:, 4:15], pred[:, :, 4:15]))
    assert _close(sdb.fde(tc, pc, t0=0, t=10), oc.fde(target[:, :10], pred[:, :, :10]))
    assert _close(sdb.apd(pc, t0=5), oc.apd(pred[:, :, 5:]), 1e-4)
    assert sdb.apd(pc[:0]).shape == (0,)                                   # empty batch of windows
    with pytest.raises(NativeError):
        sdb.apd(pred)                                                      # no CPU path
    with pytest.raises(NativeError):
        sdb.apd(torch.zeros(1, 92, 2, 3, 3, device=cuda_device))           # more than 91 samples
    with pytest.raises(NotImplementedError):
        sdb.ade(tc, pc, reduction="none")


def test_metrics_properties_at_bench_shape(cuda_device):
    import skeletondiffusion_b200 as sdb
    W, S, T, F = 512, 50, 120, 63
    g = torch.Generator(device=cuda_device).manual_seed(3)
    pred = torch.rand(W, S, T, 21, 3, device=cuda_device, generator=g) * 2 - 1
    target = torch.rand(W, T, 21, 3, device=cuda_device, generator=g) * 2 - 1
    a, f, d = sdb.motion_metrics(target, pred)
    # against plain torch in float64 on the same device, chunked
    diff = (pred.reshape(W, S, T, F).double() - target.reshape(W, 1, T, F).double()).norm(dim=-1)
    assert _close(a, diff.mean(-1).min(-1).values) and _close(f, diff[..., -1].min(-1).values)
    iu = torch.triu_indices(S, S, offset=1, device=cuda_device)
    for w0 in range(0, W, 64):
        p64 = pred[w0:w0 + 64].reshape(-1, S, T * F).double()
        assert _close(d[w0:w0 + 64], (p64[:, iu[0]] - p64[:, iu[1]]).norm(dim=-1).mean(-1))
    # linear in the pose scale; APD invariant to a permutation of the samples and to a common translation
    a2, f2, d2 = sdb.motion_metrics(target, pred, scale=1.7)
    assert _close(a2, a * 1.7, 1e-6) and _close(f2, f * 1.7, 1e-6) and _close(d2, d * 1.7, 1e-6)
    perm = torch.randperm(S, device=cuda_device, generator=g)
    a3, f3, d3 = sdb.motion_metrics(target, pred[:, perm])
    assert torch.equal(a3, a) and torch.equal(f3, f) and _close(d3, d, 1e-6)
    # a sample equal to the target gives ADE = FDE = 0; identical samples give APD = 0
    pred[:, 7] = target
    a4, f4, _ = sdb.motion_metrics(target, pred)
    assert float(a4.abs().max()) == 0.0 and float(f4.abs().max()) == 0.0
    same = target[:, None].expand(W, 4, T, 21, 3).contiguous()
    assert float(sdb.apd(same).abs().max()) == 0.0


def test_multimodal_metrics_vs_oracle_and_reference_fixture(cuda_device):
    """MMADE / MMFDE with ragged ground-truth groups (1, several, many per window) against the oracle's statement of
    src/metrics/multimodal.py:105-135 and against values produced by the reference's own mmade / mmfde (tests/golden/mm_metrics.npz,
    made by tests/golden/make_mm_metrics.py)."""
    import numpy as np
    import os
    import skeletondiffusion_b200 as sdb
    from oracle import skeldiff_oracle as oc
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mm_metrics.npz"))
    pred = torch.from_numpy(z["pred"])
    counts = z["counts"].tolist()
    flat = torch.from_numpy(z["mm_gt"])
    mm_gt, o = [], 0
    for c in counts:
        mm_gt.append(flat[o:o + c])
        o += c
    a, f = sdb.multimodal_metrics(pred.to(cuda_device), [g.to(cuda_device) for g in mm_gt], scale=1.0)
    assert torch.allclose(a.cpu(), oc.mmade(pred, mm_gt), rtol=1e-5, atol=1e-6)
    assert torch.allclose(f.cpu(), oc.mmfde(pred, mm_gt), rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.cpu(), torch.from_numpy(z["mmade"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(f.cpu(), torch.from_numpy(z["mmfde"]), rtol=1e-5, atol=1e-6)
    # frame window t0 / t and the pose-box scale
    a2, f2 = sdb.multimodal_metrics(pred.to(cuda_device), [g.to(cuda_device) for g in mm_gt], scale=1.5, t0=2, t=9)
    ref_a = 1.5 * oc.mmade(pred[:, :, 2:9], [g[:, 2:9] for g in mm_gt])
    assert torch.allclose(a2.cpu(), ref_a, rtol=1e-5, atol=1e-6)
    # twice the same call: bitwise
    a3, f3 = sdb.multimodal_metrics(pred.to(cuda_device), [g.to(cuda_device) for g in mm_gt], scale=1.0)
    assert torch.equal(a3, a) and torch.equal(f3, f)


def test_long_term_prediction_matches_the_reference_function(cuda_device):
    """long_term_prediction_best_every50 (src/eval_utils.py:44-67) around a deterministic stand-in predictor: the reference's own
    function produced tests/golden/long_term.npz (make_long_term.py); ours runs the selection / feedback loop on the device."""
    import numpy as np
    import os
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import fake_predictor
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long_term.npz"))
    S, T, factor = int(z["S"]), int(z["T"]), float(z["factor"])
    spec = sdb.get_skeleton("h36m")
    predict = fake_predictor(int(z["seed"]), S, T)
    d = cuda_device
    tgt, pred, obs = sdb.long_term_prediction_best_every50(torch.from_numpy(z["data"]).to(d), torch.from_numpy(z["target"]).to(d), None, spec,
                                                           num_samples=S, pred_length=T, long_term_factor=factor, predict_fn=predict)
    assert tuple(pred.shape) == tuple(z["out_pred"].shape)
    assert torch.allclose(pred.cpu(), torch.from_numpy(z["out_pred"]), atol=2e-6, rtol=1e-5)
    assert torch.allclose(tgt.cpu(), torch.from_numpy(z["out_target"]), atol=1e-6)
    assert torch.allclose(obs.cpu(), torch.from_numpy(z["out_obs"]), atol=1e-6)


def test_best_sample_vs_oracle(cuda_device):
    import skeletondiffusion_b200 as sdb
    g = torch.Generator().manual_seed(4)
    W, S, T, J = 9, 50, 13, 21
    pred = torch.rand(W, S, T, J, 3, generator=g) * 2 - 1
    target = torch.rand(W, T, J, 3, generator=g) * 2 - 1
    ref_idx = torch.linalg.norm(pred - target.unsqueeze(1), dim=-1).mean(-1).mean(-1).min(dim=-1).indices      # src/metrics/utils.py:24
    best, tail, idx = sdb.best_sample(pred.to(cuda_device), target.to(cuda_device), keep_frames=5, scale=1.5)
    assert torch.equal(idx.cpu().long(), ref_idx)
    ref_best = pred[torch.arange(W), ref_idx] * 1.5
    assert torch.equal(best.cpu(), ref_best) and torch.equal(tail.cpu(), ref_best[:, -5:])
