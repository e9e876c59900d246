"""Prints the relative error (max|a-b|/max|b|) of the final latents vs the reference goldens for each precision path."""
import sys, torch
sys.path.insert(0, '.')
from tests import _golden as G
d = torch.device('cuda:0')
for name in G.DATASET_CASES:
    case = G.load_npz(name)
    row = []
    for prec in ("fp32", "bf16x3", "bf16"):
        spec, ae, diff, _, _ = G.dataset_models(case, device=d, precision=prec)
        W, S = int(case["windows"]), int(case["samples"])
        lat, _ = diff.sample(batch_size=W * S, x_cond=case["z_past"].to(d), start_noise=case["start_noise"].to(d), sampling_noise=case["sampling_noise"].to(d))
        row.append(f"{prec} {G.rel_err(lat.cpu(), case['latents']):.2e}")
    print(name, " | ".join(row))
