#!/usr/bin/env python
"""Throughput / latency lines for the BASELINE.json configurations that bench.py's headline does not cover
(bench.py measures configs[1], the AMASS eval batch).  One JSON line per measurement on stdout:

  config 1  README plug-and-play (Denoiser(dim=96, num_nodes=16), T = 10, sample(batch_size=4)): latency per call,
            eager launches and CUDA graph, next to the unmodified reference (oracle/_ref) on the host cores and on the same GPU
  config 2g the reference's own eager PyTorch path on THIS GPU for the AMASS eval configuration (GPU-vs-GPU anchor)
  config 3  AMASS shape with if_run_as_isotropic=True (U = I, Lambda = 1) through the same kernels
  config 4  Human3.6M / FreeMan: 50 samples x 4096 synthetic windows sharded over the ranks (torchrun), chunks of 512 windows
  config 5  batch / timestep sweep of the fused step kernel alone and of the sampling loop (step + Denoiser)

  python bench_configs.py [--configs 1,2g,3,4,5] [--windows 4096]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_configs.py --configs 4

All times are CUDA events on the launching stream after warm-up, max over ranks; inputs are synthetic (observations
N(0, 0.3^2) clipped to the unit box) and weights reference-style random init, as in bench.py."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402  (synthetic_obs, reference_models: same inputs and reference construction as the headline)


def cuda_time(dev, fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e-3 / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


def readme_diffusion(sdb, dev, precision):
    torch.manual_seed(0)
    N = 16
    rand = (torch.rand(N, N) >= 0.5).float()
    corr = (rand + rand.T) // 2                                   # README.md:78-80
    Sigma_N, Lambda_N, U = sdb.get_cov_from_corr(corr, if_sigma_n_scale=True, sigma_n_scale="spectral")
    model = sdb.Denoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    return sdb.NonisotropicGaussianDiffusion(Sigma_N=Sigma_N, Lambda_N=Lambda_N, U=U, model=model, timesteps=10, precision=precision).to(dev).eval()


def config1(sdb, dev, args):
    """README configuration: the latency of one diffusion.sample(batch_size=4) call (10 steps, N = 16, depth 1)."""
    out = {}
    for prec in ("fp16x2", "fp32"):
        diff = readme_diffusion(sdb, dev, prec)
        for graph in (False, True):
            diff.use_cuda_graph = graph
            t = cuda_time(dev, lambda: diff.sample(batch_size=4), iters=50, warm=5)
            t0 = time.perf_counter()
            for _ in range(50):
                diff.sample(batch_size=4)
            torch.cuda.synchronize(dev)
            out[f"{prec}_{'graph' if graph else 'eager'}"] = {"device_ms": t * 1e3, "host_wall_ms": (time.perf_counter() - t0) / 50 * 1e3}
    line = {"config": "1: README plug-and-play, Denoiser(dim=96, cond_dim=0, num_nodes=16), T=10, sample(batch_size=4)", "unit": "ms per sample() call",
            "ours": out}
    from oracle import make_ref
    if make_ref.add_to_path():
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            from src.core.network import Denoiser as RefDenoiser
            from src.core.diffusion import NonisotropicGaussianDiffusion as RefDiff
            from src.core.diffusion.utils import get_cov_from_corr as ref_cov
            for where in ("cpu", "cuda"):
                torch.manual_seed(0)
                rand = (torch.rand(16, 16) >= 0.5).float()
                corr = (rand + rand.T) // 2
                S_, L_, U_ = ref_cov(correlation_matrix=corr, if_sigma_n_scale=True, sigma_n_scale="spectral")
                m = RefDenoiser(dim=96, cond_dim=0, out_dim=96, channels=16, num_nodes=16)
                rd = RefDiff(Sigma_N=S_, Lambda_N=L_, U=U_, model=m, timesteps=10).to(where).eval()
                with torch.no_grad():
                    for _ in range(3):
                        rd.sample(batch_size=4)
                    if where == "cuda":
                        torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    for _ in range(20):
                        rd.sample(batch_size=4)
                    if where == "cuda":
                        torch.cuda.synchronize(dev)
                line[f"reference_{where}_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        line["reference"] = f"unmodified reference (oracle/_ref) diffusion.sample(batch_size=4), torch {torch.__version__}: cpu = {os.cpu_count()} host threads, cuda = stock eager on this GPU (wall clock with synchronize)"
    emit(**line)


def config2_gpu_reference(sdb, dev, args):
    """The reference's eager PyTorch path on this GPU, AMASS eval configuration, at the largest window counts that fit."""
    from oracle import make_ref
    if not make_ref.add_to_path():
        emit(config="2g: reference eager on this GPU", unavailable="oracle/_ref is absent")
        return
    spec = sdb.get_skeleton("amass")
    ae, diff, get_pred = bench.reference_models(spec, "amass", False, dev)
    res = {}
    for w in (8, 64, 256):
        try:
            bench.reference_time(spec, ae, diff, get_pred, min(w, 8), 50, dev)
            dt = min(bench.reference_time(spec, ae, diff, get_pred, w, 50, dev) for _ in range(3))
            res[str(w)] = {"motions_per_s": w * 50 / dt, "ms": dt * 1e3}
        except torch.cuda.OutOfMemoryError:
            res[str(w)] = "out of memory"
            torch.cuda.empty_cache()
            break
    # ours on the same window counts (eager launches and the whole-pipeline graph)
    ae_o, diff_o = bench.oracle_state(spec, False)
    model = (ae_o.to(dev).eval(), diff_o.to(dev).eval())
    model[1].precision = "fp16x2"
    ours = {}
    for w in (8, 64, 256, 512):
        obs = bench.synthetic_obs(spec, w, 123).to(dev)
        g = sdb.GraphedPrediction(model, w, 50, spec.pred_length, dev)
        t = cuda_time(dev, lambda: g(obs), iters=3)
        ours[str(w)] = {"motions_per_s": w * 50 / t, "ms": t * 1e3}
        del g
    emit(config="2g: AMASS eval pipeline, the reference's own eager PyTorch code on this GPU next to ours (same GPU, same window counts)", unit="motions/s",
         reference_gpu_eager=res, ours_graph_fp16x2=ours,
         note="reference = unmodified get_prediction from oracle/_ref, .to('cuda'), wall clock around torch.cuda.synchronize; best of 3")


def pipeline_throughput(sdb, dev, spec, isotropic, windows, chunk, rank, world, precision="fp16x2", perturbed=False):
    """Whole job of `windows` windows x 50 samples: this rank's shard in chunks of `chunk` windows through the graph."""
    import torch.distributed as dist
    from skeletondiffusion_b200.distributed import gather_window_metrics
    from skeletondiffusion_b200.testing import synth_state_dict
    ae, diff = sdb.build_models(spec, "cpu", if_run_as_isotropic=isotropic)
    if perturbed:
        diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
        ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
    model = (ae.to(dev).eval(), diff.to(dev).eval())
    model[1].precision = precision
    lo, hi = sdb.shard_windows(windows, rank, world)
    mine = hi - lo
    chunk = min(chunk, mine)
    obs = bench.synthetic_obs(spec, mine, 1000 + rank).to(dev)
    tgt = bench.synthetic_obs(spec, mine, 2000 + rank)[:, :1].expand(-1, spec.pred_length, -1, -1).contiguous().to(dev)
    g = sdb.GraphedPrediction(model, chunk, 50, spec.pred_length, dev)
    g_tail = sdb.GraphedPrediction(model, mine % chunk, 50, spec.pred_length, dev) if mine % chunk else None
    out = {k: torch.empty(mine, device=dev) for k in ("ade", "fde", "apd")}

    def job():
        for c0 in range(0, mine, chunk):
            c1 = min(mine, c0 + chunk)
            p = (g if c1 - c0 == chunk else g_tail)(obs[c0:c1], window_offset=lo + c0)
            a, f, d = sdb.motion_metrics(tgt[c0:c1], p, scale=spec.pose_box_size)
            out["ade"][c0:c1], out["fde"][c0:c1], out["apd"][c0:c1] = a, f, d

    job()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = gather_window_metrics(out, windows, rank, world)
    return float(ms.item()) * 1e-3, gathered


def config3(sdb, dev, args, rank, world):
    spec = sdb.get_skeleton("amass")
    res = {}
    for iso in (True, False):
        t, _ = pipeline_throughput(sdb, dev, spec, iso, 512 * world, 512, rank, world)
        res["isotropic" if iso else "nonisotropic"] = {"motions_per_s": 512 * world * 50 / t, "ms_per_512_windows": t * 1e3}
    if rank == 0:
        emit(config="3: AMASS shape, if_run_as_isotropic=True (U = I, Lambda = 1, Sigma = 0) through the same kernels, next to the nonisotropic run",
             unit="motions/s", n_gpus=world, precision="fp16x2", **res)


def config4(sdb, dev, args, rank, world):
    for name in ("h36m", "freeman"):
        spec = sdb.get_skeleton(name)
        for perturbed in (False, True):
            t, gathered = pipeline_throughput(sdb, dev, spec, False, args.windows, 512, rank, world, perturbed=perturbed)
            if rank == 0:
                emit(config=f"4: {name} ({spec.num_nodes} nodes, obs {spec.obs_length} -> pred {spec.pred_length} frames), {args.windows} windows x 50 samples "
                            f"sharded over {world} GPU(s), chunks of 512 windows, ADE/FDE/APD per window + final NCCL gather inside the timed job",
                     unit="motions/s", n_gpus=world, precision="fp16x2",
                     weights="dense perturbed graph influence" if perturbed else "reference-style random init (identity graph influence)",
                     motions_per_s=args.windows * 50 / t, job_s=t, gathered_windows=int(gathered["ade"].numel()),
                     mean_ade=float(gathered["ade"].mean()), mean_apd=float(gathered["apd"].mean()))


def config5(sdb, dev, args):
    """Sweep: the fused reverse-step kernel alone (HBM roofline fraction per batch) and the sampling loop (T steps of Denoiser + step)."""
    spec = sdb.get_skeleton("amass")
    N, D = spec.num_nodes, 96
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    step = {}
    _, diff10 = sdb.build_models(spec, dev, diffusion_timesteps=10)
    for B in (1, 16, 256, 4096, 25600, 65536):
        x_t, x0, eps = (torch.randn(B, N, D, device=dev) for _ in range(3))
        t = bench._timed_kernel(dev, lambda: diff10._reverse_step(x_t, x0, eps, 5), iters=20)
        by = 4.0 * B * N * D * 4
        step[str(B)] = {"us": t * 1e6, "GB_per_s": by / t / 1e9, "frac_of_hbm_peak": by / t / 1e9 / hbm}
        del x_t, x0, eps
    emit(config="5a: fused reverse-step kernel alone, AMASS N=21, batch sweep (the time of one step does not depend on T)", unit="us per launch",
         hbm_peak_gbs=hbm, bytes_per_sample_step=4 * N * D * 4, step_kernel=step)
    loop = {}
    for T, batches in ((10, (1, 16, 256, 4096, 25600, 65536)), (100, (1, 256, 4096)), (1000, (1, 256))):
        _, diff = sdb.build_models(spec, dev, diffusion_timesteps=T)
        diff.precision = "fp16x2"
        for B in batches:
            cond = torch.tanh(torch.randn(B, N, D, device=dev))
            for graph in ((True, False) if B <= 256 else (False,)):
                diff.use_cuda_graph = graph
                t = cuda_time(dev, lambda: diff.sample(batch_size=B, x_cond=cond), iters=2 if B * T > 100000 else 5, warm=2)
                loop[f"T={T},B={B},{'graph' if graph else 'eager'}"] = {"ms": t * 1e3, "latents_per_s": B / t, "us_per_step": t / T * 1e6}
            del cond
        del diff
        torch.cuda.empty_cache()
    emit(config="5b: sampling loop (T x (Denoiser + fused step)), AMASS N=21, fp16x2, batch / timestep sweep", unit="ms per sample() call", loop=loop)


def config_train(sdb, dev, args):
    """Training step of the diffusion: TrainerDiffusion.loss with k = 50 samples per observation
    (src/core/trainer.py:224-234, similarity in latent space) + backward, AMASS configuration, 64 observations per step.
    Ours: all 3 200 loss values from the inference kernels, backward through the 64 selected rows only (training.SparseRowLoss).
    Reference: its own p_losses + autograd on this GPU (oracle/_ref)."""
    from skeletondiffusion_b200 import training
    spec = sdb.get_skeleton("amass")
    N, B, k = spec.num_nodes, 64, 50
    g = torch.Generator().manual_seed(1)
    x_start = torch.tanh(torch.randn(B, N, 96, generator=g)).to(dev)
    x_cond = torch.tanh(torch.randn(B, N, 96, generator=g)).to(dev)
    line = {"config": f"train: AMASS diffusion training step, {B} observations x k = {k} samples (best-of-k relaxation), loss + backward", "unit": "ms per step"}
    _, diff = sdb.build_models(spec, "cpu")
    with torch.enable_grad():
        for prec in ("fp32", "fp16x2"):
            diff = diff.to(dev).train()
            diff.precision = prec

            def step():
                diff.zero_grad(set_to_none=True)
                loss, w, _ = diff(x_start, x_cond=x_cond, n_train_samples=k)
                sim, _ = training.ksimilarity_loss(loss, B)
                (sim * w).mean().backward()

            line[f"ours_{prec}_ms"] = cuda_time(dev, step, iters=10, warm=3) * 1e3

        def step_dense():
            diff.zero_grad(set_to_none=True)
            loss, w, _ = diff(x_start.repeat_interleave(k, 0), x_cond=x_cond.repeat_interleave(k, 0), n_train_samples=1)
            sim, _ = training.ksimilarity_loss(loss, B)
            (sim * w[::k]).mean().backward()

        line["ours_dense_autograd_all_rows_ms"] = cuda_time(dev, step_dense, iters=5, warm=2) * 1e3
        from oracle import make_ref
        if make_ref.add_to_path():
            ae_r, diff_r, _ = bench.reference_models(spec, "amass", False, dev)
            diff_r = diff_r.train()

            def ref_step():
                diff_r.zero_grad(set_to_none=True)
                loss, w, _ = diff_r(x_start, x_cond=x_cond, n_train_samples=k)
                idx = loss.detach().view(B, -1).min(-1).indices
                (torch.gather(loss.view(B, -1), 1, idx.unsqueeze(1)).squeeze(-1) * w).mean().backward()

            line["reference_gpu_eager_ms"] = cuda_time(dev, ref_step, iters=5, warm=2) * 1e3
            line["reference"] = "unmodified reference p_losses + torch autograd (oracle/_ref), stock eager PyTorch on this GPU"
    emit(**line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2g,3,4,5,train")
    ap.add_argument("--windows", type=int, default=4096)
    args = ap.parse_args()
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.distributed import init_from_env
    rank, local_rank, world = init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    which = args.configs.split(",")
    with torch.no_grad():
        if "1" in which and rank == 0:
            config1(sdb, dev, args)
        if "2g" in which and rank == 0:
            config2_gpu_reference(sdb, dev, args)
        if "3" in which:
            config3(sdb, dev, args, rank, world)
        if "4" in which:
            config4(sdb, dev, args, rank, world)
        if "5" in which and rank == 0:
            config5(sdb, dev, args)
        if "train" in which and rank == 0:
            config_train(sdb, dev, args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
