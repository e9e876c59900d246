"""CUDA-event timing of motion_metrics_kernel at the bench shape (512 windows x 50 samples x 120 frames x 63 features)."""
import json, sys, torch
import skeletondiffusion_b200 as sdb
W, S, T, F = 512, 50, 120, 63
dev = torch.device("cuda:0")
pred = torch.rand(W, S, T, 21, 3, device=dev) * 2 - 1
target = torch.rand(W, T, 21, 3, device=dev) * 2 - 1
for _ in range(5): sdb.motion_metrics(target, pred, 1.5)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): sdb.motion_metrics(target, pred, 1.5)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
alg = 4 * (S + 1) * T * F * W + 12 * W
print(json.dumps({"kernel": "motion_metrics_kernel", "shape": [W, S, T, F], "ms": ms, "algorithmic_bytes": alg,
                  "achieved_GBps": alg / ms / 1e6, "note": "input 787 MB > 126 MB L2"}))
