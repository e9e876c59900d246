#!/usr/bin/env python
"""Headline benchmark: predicted motions/s of the full sampling pipeline (encode past -> T-step
nonisotropic reverse diffusion with the Denoiser -> decode), AMASS eval configuration
(BASELINE.json configs[1]: 512 observed windows x 50 samples = 25 600 motions per step), synthetic
observations, reference-style random-init weights.

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            # the UNMODIFIED reference (oracle/_ref) on the host cores; oracle port if absent

Rank 0 prints ONE JSON line (see README/DESIGN for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "predicted motions/sec (obs x samples, full sampling loop: encode -> 10-step nonisotropic diffusion -> decode)"
UNIT = "motions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dataset", default="amass")
    ap.add_argument("--windows", type=int, default=512, help="observed windows per GPU per step (configs/config_eval/config.yaml:27)")
    ap.add_argument("--samples", type=int, default=50)
    ap.add_argument("--precision", default=os.environ.get("SKELDIFF_PRECISION", "fp16x2"), choices=["fp32", "bf16", "bf16x3", "fp16x2"],
                    help="headline path: fp16x2 (default) / bf16x3 = fp32-grade tensor-core paths (two fp16 / three bf16 operand planes), "
                         "fp32 = FFMA2, bf16 = bf16 activations")
    ap.add_argument("--cpu-windows", type=int, default=64, help="windows of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--perturbed", action="store_true", help="dense non-identity graph-influence matrices (trained-model-like)")
    ap.add_argument("--no-graph", action="store_true", help="eager kernel launches instead of the whole-pipeline CUDA graph")
    ap.add_argument("--no-general", action="store_true", help="skip the second headline (dense graph-influence weights)")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary precisions (shorter run, e.g. under ncu)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the reference arm (host cores); cuda = the same stock eager code on the GPU")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        mhz = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def synthetic_obs(spec, windows, seed):
    g = torch.Generator().manual_seed(seed)
    # root-relative joints ~ N(0, 0.3^2) clipped to the unit box
    return (torch.randn(windows, spec.obs_length, spec.num_nodes, 3, generator=g) * 0.3).clamp_(-1, 1)


def oracle_state(spec, perturbed, seed=0):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    ae, diff = sdb.build_models(spec, "cpu", seed=seed)
    if perturbed:
        diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
        ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
    return ae, diff


def cpu_reference_time(spec, ae, diff, windows, samples, repeats=1):
    """The reference algorithm's CPU path (oracle port, fp32 PyTorch, all host threads) on `windows` windows."""
    from oracle import skeldiff_oracle as oc
    ae_sd = {k: v.detach().cpu() for k, v in ae.state_dict().items()}
    diff_sd = {k: v.detach().cpu() for k, v in diff.state_dict().items()}
    tab = {k: v for k, v in diff_sd.items() if not k.startswith("model.")}
    cfg = dict(dim=96, cond_dim=96, depth=4, attn_heads=8, attn_dim_head=32, node_types=spec.nodes_type_id, learn_influence=True,
               enc_num_layers=spec.enc_num_layers)
    obs = synthetic_obs(spec, windows, 123)
    B = windows * samples
    g = torch.Generator().manual_seed(5)
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            start = torch.randn(B, spec.num_nodes, 96, generator=g)
            noise = torch.randn(B, 9, spec.num_nodes, 96, generator=g)
            oc.get_prediction(ae_sd, diff_sd, cfg, tab, tab["U"], obs, samples, spec.pred_length, start, noise)
            best = min(best, time.perf_counter() - t0)
    return B / best, best


NUM_JOINTS = {"amass": 22, "h36m": 17, "freeman": 18}


def reference_models(spec, dataset, perturbed, device):
    """The UNMODIFIED reference classes from oracle/_ref (copied there by oracle/make_ref.py), dataset configuration, reference
    initialisation under torch.manual_seed(0) (configs/config_eval).  Returns (autoencoder, diffusion, get_prediction)."""
    import contextlib
    import warnings
    warnings.filterwarnings("ignore")
    with contextlib.redirect_stdout(sys.stderr):      # the reference prints while it builds its modules; stdout carries the JSON line only
        return _reference_models(spec, dataset, perturbed, device)


def _reference_models(spec, dataset, perturbed, device):
    from src.core import AutoEncoder as RefAutoEncoder, DiffusionManager as RefDiffusionManager
    from src.data.skeleton import create_skeleton
    from src.eval_prepare_model import get_prediction as ref_get_prediction
    torch.manual_seed(0)
    sk = create_skeleton(dataset_name=dataset, motion_repr_type="SkeletonRescalePose", num_joints=NUM_JOINTS[dataset], if_consider_hip=False,
                         obs_length=spec.obs_length, pred_length=spec.pred_length, pose_box_size=spec.pose_box_size)
    arch = dict(depth=4, attn_heads=8, attn_dim_head=32, use_attention=True, self_condition=False, norm_type="none", learn_influence=True)
    mgr = RefDiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=sk, covariance_matrix_type="adjacency",
                              num_nodes=sk.num_nodes, node_types=sk.nodes_type_id, diffusion_conditioning=True, latent_size=96,
                              diffusion_timesteps=10, diffusion_objective="pred_x0", beta_schedule="cosine", diffusion_arch=arch)
    diff = mgr.get_diffusion()
    ae = RefAutoEncoder(num_nodes=sk.num_nodes, encoder_hidden_size=96, decoder_hidden_size=96, latent_size=96, node_types=sk.nodes_type_id,
                        input_size=3, z_activation="tanh", enc_num_layers=spec.enc_num_layers, recurrent_arch_enc="StaticGraphGRU",
                        recurrent_arch_decoder="StaticGraphGRU", output_size=3, if_consider_hip=False)
    if perturbed:
        from skeletondiffusion_b200.testing import synth_state_dict
        diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
        ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
    return ae.to(device).eval(), diff.to(device).eval(), ref_get_prediction


def reference_time(spec, ae, diff, get_pred, windows, samples, device):
    """One timed call of the reference's own get_prediction (src/eval_prepare_model.py:118-121) on `windows` windows."""
    obs = synthetic_obs(spec, windows, 123).to(device)
    with torch.no_grad():
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pred = get_pred(obs, (ae, diff), num_samples=samples, pred_length=spec.pred_length, diffusion_conditioning=True)
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
    assert tuple(pred.shape) == (windows, samples, spec.pred_length, spec.num_nodes, 3)
    return dt


def run_reference(args):
    """Reference arm: the reference's own implementation of the path on the host cores (all threads), a bounded sample of the
    workload per step.  oracle/_ref (the unmodified reference package, see oracle/make_ref.py) when present -> kind "reference";
    otherwise the oracle port -> kind "port".  --ref-device cuda times the same stock eager PyTorch code on the GPU instead
    (the like-for-like GPU comparison of SURVEY 8d; reported under its own key, never as the CPU baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import skeletondiffusion_b200 as sdb
    from oracle import make_ref
    spec = sdb.get_skeleton(args.dataset)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = max(1, args.cpu_windows)
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 1))
    have_ref = make_ref.add_to_path()
    on_gpu = args.ref_device == "cuda"
    if on_gpu and not have_ref:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is absent: run python oracle/make_ref.py in the build container"}))
        return
    if have_ref:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if on_gpu else torch.device("cpu")
        ae, diff, get_pred = reference_models(spec, args.dataset, args.perturbed, device)
        for _ in range(warm):
            reference_time(spec, ae, diff, get_pred, 1 if not on_gpu else w, args.samples, device)
        times = [reference_time(spec, ae, diff, get_pred, w, args.samples, device) for _ in range(min(steps, 5 if on_gpu else 3))]
        kind = "reference"
        what = f"{w} windows x {args.samples} samples, unmodified reference get_prediction (oracle/_ref, src/eval_prepare_model.py:118-121), torch {torch.__version__} {'CUDA eager' if on_gpu else 'CPU'} fp32"
    else:
        ae, diff = oracle_state(spec, args.perturbed)
        cpu_reference_time(spec, ae, diff, 1, args.samples)       # warm-up (thread pools, allocator)
        times = [cpu_reference_time(spec, ae, diff, w, args.samples)[1] for _ in range(min(steps, 3))]
        kind = "port"
        what = f"{w} windows x {args.samples} samples, oracle/skeldiff_oracle.py get_prediction, torch {torch.__version__} CPU fp32"
    dt = sum(times) / len(times)
    value = w * args.samples / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic observations, random-init weights" + (" with dense perturbed graph-influence matrices" if args.perturbed else ""),
            "config": {"workload": f"{args.dataset} eval: {w} windows x {args.samples} samples per step (bounded sample of the 512-window batch)",
                       "windows_per_step": w, "samples": args.samples, "timesteps": 10, "pred_length": spec.pred_length,
                       "device": "cuda (stock eager PyTorch)" if on_gpu else "cpu"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if on_gpu:
        line["reference_gpu_eager"] = {"value": value, "unit": UNIT, "kind": kind, "sample": what, "device": torch.cuda.get_device_name(0)}
    else:
        line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": what}
    print(json.dumps(line))


def _timed_kernel(dev, fn, iters=10):
    """Average device time of one launch: `iters` launches between two CUDA events on the launching stream.  A device-side
    sleep is queued first so that the host (ctypes call + argument marshalling, 0.1-0.4 ms per call) runs ahead of the GPU
    and the launches execute back to back; without it a 0.2 ms kernel was measured as the host's enqueue time."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    torch.cuda._sleep(int(4e7))              # ~20 ms of GPU idle-spin ahead of the timed region
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize(dev)
    return sum(ev[i].elapsed_time(ev[i + 1]) for i in range(iters)) / iters * 1e-3


def kernel_rooflines(dev, spec, diff, peaks, precision="fp16x2"):
    """The kernels that dominate each path, timed alone with CUDA events on torch's current stream (the stream the
    library launches on) at the full batch B = 25 600; every operand set is larger than the 126 MB L2.
    Algorithmic work per launch follows DESIGN.md section 4 / SURVEY section 8d."""
    from skeletondiffusion_b200 import _native as nv
    lib = nv.load()
    B, N, C = 25600, spec.num_nodes, 192
    plan = diff.model.layers[0][0].block2.proj.plan()
    x = torch.randn(B, N, C, device=dev)
    res = torch.randn(B, N, C, device=dev)
    out = torch.empty(B, N, C, device=dev)
    ss = torch.zeros(1, 2 * C, device=dev)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    src = "MEASURED_PEAKS.json" if peaks.get("_measured") else "fallback"
    hbm, tfl = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0)
    flops = 2.0 * B * N * C * C                      # algorithmic fp32 FLOPs of one 192->192 graph-linear launch
    rl = {}
    # (1) fp32-grade graph-linear on the tensor cores, fp32 activations in HBM: fp16x2 = 3 fp16 MMAs per fp32 product (two operand
    # planes), bf16x3 = 6 bf16 MMAs (three planes).  Floors for this launch: HBM 3 x 413 MB / peak = 0.19 ms; tensor 3 (6) x
    # 39.6 GFLOP / peak = 0.07 (0.14) ms  =>  HBM-bound.  The layer with the Denoiser's commonest epilogue (tanh + residual).
    by3 = B * N * C * 4 * 3.0                        # read activations + read residual + write output, fp32
    for key, prec, mmas in (("glin_tc3", "fp16x2", 3), ("glin_tc3_bf16x3", "bf16x3", 6)):
        t = _timed_kernel(dev, lambda: plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, out=out, precision=prec))
        rl[key] = {"kernel": f"glin_tc3_kernel<PL={2 if mmas == 3 else 3}> (tcgen05, {prec} operand split): graph-linear 192->192 +scale/shift +tanh +residual, fp32 I/O, B=25600",
                   "bound": "hbm", "achieved": by3 / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by3 / t / 1e9 / hbm,
                   "traffic": traffic.get("glin_tc3_kernel_pl2" if mmas == 3 else "glin_tc3_kernel"), "ms": t * 1e3, "bytes_per_sample_layer": N * C * 4 * 3,
                   "tensor": {"algorithmic_tflops": flops / t / 1e12, "issued_tflops_16bit": mmas * flops / t / 1e12,
                              "issued_frac_of_bf16_peak": mmas * flops / t / 1e12 / tfl, "peak_tflops": tfl},
                   "flops_per_sample_layer": 2.0 * N * C * C, "peak_source": src}
    # the widest layer (to_qkv 192 -> 768, 20 % of a step): tensor-bound, reported against the dense 16-bit peak
    att = diff.model.layers[0][1]                 # Residual(PreNorm(Attention))
    lq = att.fn.fn.to_qkv.plan(fold_gain=att.fn.norm.g)
    oq = torch.empty(B, N, 768, device=dev)
    tq = _timed_kernel(dev, lambda: lq.forward(x, out=oq, precision=precision if precision in nv.FP32_GRADE_TC else "fp16x2"))
    fq = 2.0 * B * N * C * 768
    mm = 6 if precision == "bf16x3" else 3
    rl["to_qkv"] = {"kernel": "glin_tc3_kernel: to_qkv 192->768 (activation-stationary schedule), fp32 I/O, B=25600", "bound": "tensor",
                    "achieved": mm * fq / tq / 1e12, "peak": tfl, "unit": "TFLOP/s", "frac": mm * fq / tq / 1e12 / tfl, "traffic": None, "ms": tq * 1e3,
                    "algorithmic_tflops": fq / tq / 1e12, "mmas_per_product": mm, "hbm_gbs": B * N * (C + 768) * 4.0 / tq / 1e9, "peak_source": src}
    del oq
    # (2) exact-fp32 FFMA2 graph-linear
    t = _timed_kernel(dev, lambda: plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, out=out, precision="fp32"))
    rl["glin_ffma2"] = {"kernel": "glin_gemm_f2_kernel (FFMA2): same layer, exact fp32", "bound": "fp32 pipe (no tensor cores)",
                        "achieved": flops / t / 1e12, "peak": 72.0, "unit": "TFLOP/s", "frac": flops / t / 1e12 / 72.0,
                        "traffic": traffic.get("glin_gemm_f2_kernel"), "ms": t * 1e3,
                        "peak_source": "scratch/ffma_peak.cu on this pool's B200: 71.6 TFLOP/s FFMA, 73.8 FFMA2"}
    # (3) bf16 graph-linear on its native bf16 tensors: K = 192 -> 96 FLOP/B, below the ridge => HBM-bound
    x16, r16, o16 = x.to(torch.bfloat16), res.to(torch.bfloat16), torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
    st = nv.stream_ptr(dev)
    scr16 = None if plan.identity else torch.empty(B, N, C, device=dev)      # raw products of a layer with a dense graph influence
    t = _timed_kernel(dev, lambda: nv.check(lib.sd_glin_forward_bf16(plan.handle, x16.data_ptr(), None, ss.data_ptr(), nv.ACT_TANH, r16.data_ptr(),
                                                                     o16.data_ptr(), 0, nv.dptr(scr16), B, st), "sd_glin_forward_bf16"))
    del scr16
    by = B * N * C * 2 * 3.0
    rl["glin_tc_bf16"] = {"kernel": "glin_tc_kernel (tcgen05/TMEM/TMA): same layer, bf16 activations", "bound": "hbm", "achieved": by / t / 1e9,
                          "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm, "traffic": traffic.get("glin_tc_kernel"), "ms": t * 1e3,
                          "tflops": flops / t / 1e12, "bytes_per_sample_layer": N * C * 2 * 3, "peak_source": src}
    # (4) fused reverse step: 3 reads + 1 write of [B, N, 96] fp32 (injected noise)
    x_t, x0, eps = (torch.randn(B, N, 96, device=dev) for _ in range(3))
    t = _timed_kernel(dev, lambda: diff._reverse_step(x_t, x0, eps, 5))
    by = 4.0 * B * N * 96 * 4
    rl["reverse_step"] = {"kernel": "reverse_step_kernel<%d> (fused nonisotropic step, FFMA2), B=25600" % N, "bound": "hbm", "achieved": by / t / 1e9,
                          "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm, "traffic": traffic.get("reverse_step_kernel"), "ms": t * 1e3,
                          "bytes_per_sample_step": 4 * N * 96 * 4, "peak_source": src}
    # (5) node attention (fp32, 8 heads x 32): reads q|k|v (768 floats) and writes 256 floats per (sample, node) row
    qkv_t = torch.randn(B, N, 768, device=dev)
    att_o = torch.empty(B, N, 256, device=dev)
    t = _timed_kernel(dev, lambda: nv.check(lib.sd_node_attention(qkv_t.data_ptr(), att_o.data_ptr(), B, N, 8, 32, st), "sd_node_attention"))
    by = B * N * (768 + 256) * 4.0
    rl["node_attention"] = {"kernel": "node_attention_bulk_kernel<%d> (cp.async.bulk ring, FFMA2), fp32 I/O, B=25600" % N, "bound": "hbm",
                            "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                            "traffic": traffic.get("node_attention_bulk_kernel"), "ms": t * 1e3, "bytes_per_sample": N * (768 + 256) * 4, "peak_source": src}
    del qkv_t, att_o, x, res, out
    # (6) evaluation metrics right after the path (ADE / FDE / APD per window): reads the 512 x 50 predictions and the targets once
    W, S, T, F = 512, 50, 120, N * 3
    pred_m = torch.rand(W, S, T, F, device=dev) * 2 - 1
    tgt_m = torch.rand(W, T, F, device=dev) * 2 - 1
    m_out = torch.empty(3, W, device=dev)
    t = _timed_kernel(dev, lambda: nv.check(lib.sd_motion_metrics(pred_m.data_ptr(), tgt_m.data_ptr(), W, S, T, F, 1.5, m_out[0].data_ptr(),
                                                                  m_out[1].data_ptr(), m_out[2].data_ptr(), st), "sd_motion_metrics"))
    by = 4.0 * (S + 1) * T * F * W + 12 * W
    rl["motion_metrics"] = {"kernel": "motion_metrics_kernel (ADE/FDE/APD, 5x5 register tiles of sample pairs), 512 windows x 50 samples", "bound": "hbm",
                            "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm,
                            "traffic": traffic.get("motion_metrics_kernel"), "ms": t * 1e3, "bytes_per_window": 4 * (S + 1) * T * F + 12,
                            "note": "FP32-issue bound, not HBM bound (DESIGN.md 4.7)", "peak_source": src}
    return rl


def decoder_frame_roofline(dev, spec, ae, obs_dev, W, S, peaks, precision, dense):
    """One decoder frame at the bench batch: (decode of 24 frames - decode of 4 frames) / 20, CUDA events on the launching stream.
    Identity influence under fp16x2: the fused tcgen05 GRU step (recurrent product + gates) + the output head;
    algorithmic bytes per (sample, node) row: h operand 96 + x-side gate products 288 + new h 96 (step) + h 96 + pose 3 (head) floats.
    Dense influence: recurrent product + per-sample gate kernel + head: h 96 + hr 288 written + hr, xr 576 + h 96 read
    + h 96 written + head 99."""
    B, N = W * S, spec.num_nodes
    lat = torch.tanh(torch.randn(B, N, 96, device=dev))
    prec = "bf16x3" if precision == "bf16" else precision

    def dec(frames):
        return _timed_kernel(dev, lambda: ae.decode(obs_dev, lat, None, ph=frames, precision=prec), iters=3)

    t = (dec(24) - dec(4)) / 20.0
    floats = (96 + 288 + 96 + 96 + 3) if not dense and prec == "fp16x2" else (96 + 288 + 576 + 96 + 96 + 99)
    by = 4.0 * floats * B * N
    hbm = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    if not dense and prec == "fp16x2":      # the fused step + the head, from their ncu --set full captures (profiles/r2_final_*.summary.txt)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("decoder_frame_identity_fp16x2")
        except Exception:
            pass
    return {"kernel": ("glin_tc3_kernel<T3_ACT_GRU> (h W_hh^T on tcgen05 + GRU gates in the epilogue) + gru_head_tiled_kernel" if not dense and prec == "fp16x2"
                       else "glin_tc3 recurrent product + gate kernel + gru_head_tiled_kernel") + ", one decoder frame, B=25600",
            "bound": "hbm", "achieved": by / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / t / 1e9 / hbm, "traffic": traffic, "ms": t * 1e3,
            "bytes_per_row_frame": 4 * floats, "peak_source": "MEASURED_PEAKS.json" if peaks.get("_measured") else "fallback"}


def run_ours(args):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200 import _native as nv
    from skeletondiffusion_b200.distributed import init_from_env, gather_window_metrics
    import torch.distributed as dist
    rank, local_rank, world = init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = nv.load()
    if not lib.sd_device_supported(local_rank):
        raise SystemExit("bench.py: device is not sm_100 (B200); this library has no other code path")
    spec = sdb.get_skeleton(args.dataset)
    ae_cpu, diff_cpu = oracle_state(spec, args.perturbed)
    ae, diff = ae_cpu.to(dev).eval(), diff_cpu.to(dev).eval()
    W, S, ph = args.windows, args.samples, spec.pred_length
    B = W * S
    obs_host = synthetic_obs(spec, W, 1000 + rank).pin_memory()
    obs_dev = obs_host.to(dev)
    pred_host = torch.empty(W, S, ph, spec.num_nodes, 3).pin_memory()
    model = (ae, diff)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["_measured"] = True
    except Exception:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}

    # The measured path: get_prediction for the fixed (windows, samples, frames) bucket replayed as ONE CUDA graph
    # (encode + 10-step sampling loop + 120-frame decode, ~1 100 launches; sdb.GraphedPrediction).  --no-graph: eager launches.
    graphed = None if args.no_graph else sdb.GraphedPrediction(model, W, S, ph, dev)

    def predict(o):
        if graphed is not None:
            return graphed(o, window_offset=rank * W)       # noise indexed by the global window: independent of the rank count
        return sdb.get_prediction(o, model, num_samples=S, pred_length=ph, diffusion_conditioning=True)

    def step_resident():
        return predict(obs_dev)

    copy_stream = torch.cuda.Stream(dev)
    pred_stage = [torch.empty(W, S, ph, spec.num_nodes, 3, device=dev) for _ in range(2)]
    e2e_state = {"i": 0, "ev": [None, None]}

    def step_e2e():
        # host -> device copy of this step's observations, prediction, device -> host read of the result.  The D2H of step k
        # runs on a copy stream while step k + 1 computes (double-buffered device staging; the timed region ends after the
        # last copy has landed), so the 774 MB/step read-back is off the critical path of every step but the last.
        i = e2e_state["i"] & 1
        o = obs_host.to(dev, non_blocking=True)
        p = predict(o)
        cur = torch.cuda.current_stream(dev)
        if e2e_state["ev"][i] is not None:
            cur.wait_event(e2e_state["ev"][i])             # the staging buffer's previous copy-out has finished
        pred_stage[i].copy_(p, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            pred_host.copy_(pred_stage[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        e2e_state["ev"][i] = ev
        e2e_state["i"] += 1
        return p

    def barrier():
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) * 1e-3

    warm = max(args.warmup, 3)
    diff.precision = args.precision
    for _ in range(warm):
        step_resident()
    with ClockSampler(local_rank) as clk:
        t_res = timed(step_resident, args.steps)
    # kernels of the library inside one step: counted on an eager (un-graphed) step -- a graph replay launches the same kernels
    launches0 = lib.sd_launch_count()
    sdb.get_prediction(obs_dev, model, num_samples=S, pred_length=ph, diffusion_conditioning=True)
    launches = (lib.sd_launch_count() - launches0) * args.steps
    step_e2e()
    t_e2e = timed(step_e2e, args.steps)
    # the same pipeline with dense (trained-checkpoint-like) graph-influence matrices on every layer: G^ != I, G_add != 0
    general = None
    if not args.perturbed and not args.no_general:
        ae_p, diff_p = oracle_state(spec, True)
        model_p = (ae_p.to(dev).eval(), diff_p.to(dev).eval())
        model_p[1].precision = args.precision
        gp = None if args.no_graph else sdb.GraphedPrediction(model_p, W, S, ph, dev)
        run_p = (lambda: gp(obs_dev)) if gp is not None else (lambda: sdb.get_prediction(obs_dev, model_p, num_samples=S, pred_length=ph, diffusion_conditioning=True))
        for _ in range(3):
            run_p()
        k = max(2, min(args.steps, 3))
        t = timed(run_p, k)
        general = {"weights": "perturbed: dense G on every graph-linear, dense G / G_add on the GRUs (what a trained checkpoint has; "
                              "skeletondiffusion_b200.testing.synth_state_dict, gain 2.5)",
                   "value": B * world * k / t, "unit": UNIT, "ms_per_step": t / k * 1e3, "steps": k,
                   "ratio_to_identity_weights": (B * world * k / t) / (B * world * args.steps / t_res)}
        del gp, model_p, ae_p, diff_p
        torch.cuda.empty_cache()
    # secondary precisions of the same pipeline (same inputs, same step definition), fewer steps
    others = {}
    eager = lambda: sdb.get_prediction(obs_dev, model, num_samples=S, pred_length=ph, diffusion_conditioning=True)
    for prec in ("bf16", "fp32", "bf16x3", "fp16x2"):
        if prec == args.precision or args.no_others:
            continue
        diff.precision = prec
        for _ in range(2):
            eager()
        k = max(2, min(args.steps, 3))
        t = timed(eager, k)
        others[prec] = {"value": B * world * k / t, "unit": UNIT, "ms_per_step": t / k * 1e3}
    diff.precision = args.precision
    # final metric exchange: per-window statistic of the predictions, gathered over NCCL (KB-class message, outside the loop)
    # (ADE / FDE / APD of the predictions against a synthetic target, sd_motion_metrics on each rank's windows)
    p = step_resident()
    tgt = synthetic_obs(spec, W, 2000 + rank)[:, :1].expand(-1, ph, -1, -1).contiguous().to(dev)
    m_ade, m_fde, m_apd = sdb.motion_metrics(tgt, p, scale=spec.pose_box_size)
    local_metric = {"ade": m_ade, "fde": m_fde, "apd": m_apd}
    gathered = gather_window_metrics(local_metric, W * world, rank, world) if world > 1 else local_metric
    if rank != 0:
        return
    motions = B * world * args.steps
    value, e2e = motions / t_res, motions / t_e2e
    rl = kernel_rooflines(dev, spec, diff, peaks, args.precision)
    try:
        if args.precision != "fp32":    # (exact fp32: the FFMA2 step kernel of round 1, DESIGN.md table in section 4)
            rl["decoder_frame"] = decoder_frame_roofline(dev, spec, ae, obs_dev, W, S, peaks, args.precision, args.perturbed)
    except Exception as e:          # a secondary line must not cost the bench line
        rl["decoder_frame"] = {"error": str(e)[:200]}
    dominant = {"fp16x2": "glin_tc3", "bf16x3": "glin_tc3_bf16x3", "fp32": "glin_ffma2", "bf16": "glin_tc_bf16"}[args.precision]
    dtype = {"fp32": "f32 (FFMA2, exact)", "bf16": "bf16 (tcgen05; stated tolerance, tests/test_gpu_tc.py)",
             "bf16x3": "f32-grade: fp32 operands split into 3 bf16 planes on tcgen05, fp32 accumulate; <=1e-4 vs the reference (tests/test_gpu_tc.py)",
             "fp16x2": "f32-grade: fp32 operands split into 2 fp16 planes (22 significand bits) on tcgen05, fp32 accumulate; <=1e-4 vs the reference on every golden (tests/test_gpu_tc.py)"}[args.precision]
    if "bf16" in others:
        others["bf16"]["note"] = "bf16 activations between layers; latents within 5e-2, ADE/FDE/APD within 2 % of the fp32 reference (tests/test_gpu_tc.py)"
    if "bf16x3" in others:
        others["bf16x3"]["note"] = "three bf16 operand planes, six MMAs per product: exact operand split, any fp32 magnitude; same <=1e-4 gate"
    if "fp32" in others:
        others["fp32"]["note"] = "exact fp32 on the FFMA2 pipe, no tensor cores"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype,
            "data": "synthetic observations (N(0,0.3^2) clipped to the unit box), reference-style random-init weights" +
                    (" with dense perturbed graph-influence matrices" if args.perturbed else ""),
            "config": {"workload": f"{args.dataset} eval config: {W} windows x {S} samples = {B} motions per GPU per step; encode({spec.obs_length} frames) -> 10-step sampling -> decode({ph} frames)",
                       "windows_per_gpu": W, "samples": S, "timesteps": 10, "pred_length": ph, "num_nodes": spec.num_nodes,
                       "parallelism": f"windows sharded over {world} GPU(s), no in-loop collective", "precision": args.precision,
                       "l2_policy": "per-step working set (activations 413 MB per tensor, noise 1.9 GB) exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": obs_host.numel() * 4 * world, "d2h_bytes_per_step": pred_host.numel() * 4 * world,
                    "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": rl[dominant],
            "roofline_other_kernels": {k: v for k, v in rl.items() if k != dominant},
            "other_precisions": others, "gathered_windows": int(gathered["ade"].numel()),
            "gathered_metrics": {k: float(v.float().mean()) for k, v in gathered.items()},
            "cuda_graph": graphed is not None}
    if general is not None:
        line["general_graph_influence"] = general
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        from oracle import make_ref
        if make_ref.add_to_path():     # the unmodified reference package (oracle/_ref), timed as itself on the host cores
            cpu = torch.device("cpu")
            r_ae, r_diff, r_pred = reference_models(spec, args.dataset, args.perturbed, cpu)
            reference_time(spec, r_ae, r_diff, r_pred, 1, S, cpu)
            dt = reference_time(spec, r_ae, r_diff, r_pred, args.cpu_windows, S, cpu)
            line["cpu_baseline"] = {"value": args.cpu_windows * S / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": f"{args.cpu_windows} windows x {S} samples ({dt:.1f} s), unmodified reference get_prediction (oracle/_ref), torch CPU fp32"}
        else:
            cpu_reference_time(spec, ae_cpu, diff_cpu, 1, S)
            v, dt = cpu_reference_time(spec, ae_cpu, diff_cpu, args.cpu_windows, S)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_windows} windows x {S} samples ({dt:.1f} s), oracle port of the reference algorithm, torch CPU fp32"}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
