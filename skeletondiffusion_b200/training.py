"""Training path of the nonisotropic latent diffusion: forward() / p_losses with autograd and the k-best-sample relaxation
(src/core/diffusion/base.py:262-307, src/core/trainer.py:205-234).

Design (B200-first, not the reference's "one autograd graph over B * k rows"):

* The loss VALUES of all B * k rows come from the inference kernels (q_sample -> Denoiser with per-sample time rows ->
  Mahalanobis loss), with no autograd state kept.
* TrainerDiffusion.loss keeps, per observation, only the sample closest to the ground truth (trainer.py:205-221): the gradient
  of the returned loss vector is non-zero in 1 row out of k.  `SparseRowLoss.backward` therefore re-runs the Denoiser WITH
  saved activations on exactly the rows whose loss gradient is non-zero and back-propagates through those (rematerialisation:
  k = 50 -> 1/50 of the backward work and of the activation memory; the parameter gradients are the same numbers, the rows
  that are dropped contribute exact zeros).
* The differentiable Denoiser is a composition of torch.autograd.Function nodes whose forward AND backward are the library's
  CUDA kernels (graph-linear: forward GEMM kernels for out and dX, sd_glin_backward_params for dW / db / dG^; Block epilogue,
  RMSNorm, node attention, loss: sd_*_forward / sd_*_backward).  torch only adds tensors, concatenates, and differentiates the
  batch-invariant pieces: the [T, .] time-conditioning table (time MLP on the T distinct steps, not on B rows) and the L1
  normalisation of the N x N influence matrices.

CUDA only, no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nv
from .plan import GlinPlan, Workspace, params_key

__all__ = ["denoiser_forward_train", "diffusion_loss_train", "SparseRowLoss", "ksimilarity_loss"]


def _gemm_precision(precision: str) -> str:
    return precision if precision in nv.FP32_GRADE_TC else "fp32"


class _TrainPlans:
    """Packed operands of one StaticGraphLinear for the training path, rebuilt when its parameters change:
    fwd = the production plan (G^ (x W^T + b)); raw = x W^T alone (for dG^); bwd = dYm W (transposed weights)."""

    def __init__(self, layer):
        self.key = params_key([layer.G, layer.weight, layer.bias])
        self.fwd = layer.plan()
        w = layer.weight.detach().float()
        w3 = w.unsqueeze(0) if w.dim() == 2 else w
        eye = torch.eye(layer.num_nodes, device=w.device)
        self.raw = GlinPlan(layer.num_nodes, layer.node_type_index, w3, None, eye)
        self.bwd = GlinPlan(layer.num_nodes, layer.node_type_index, w3.transpose(1, 2).contiguous(), None, eye)


def _plans(layer) -> _TrainPlans:
    p = getattr(layer, "_train_plans", None)
    if p is None or p.key != params_key([layer.G, layer.weight, layer.bias]):
        p = layer._train_plans = _TrainPlans(layer)
    return p


class _GraphLinearFn(torch.autograd.Function):
    """out = G^ (x W[type]^T + b[type])   (graph_structural.py:30-43)"""

    @staticmethod
    def forward(ctx, x, weight, bias, g_hat, layer, precision):
        x = x.contiguous()
        plans = _plans(layer)
        out = plans.fwd.forward(x, precision=precision)
        ctx.save_for_backward(x, weight, bias if bias is not None else x.new_empty(0), g_hat)
        ctx.layer, ctx.precision, ctx.has_bias = layer, precision, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, bias, g_hat = ctx.saved_tensors
        layer, plans, lib = ctx.layer, _plans(ctx.layer), nv.load()
        dout = dout.contiguous().float()
        B, N, OUT = dout.shape
        dev = dout.device
        st = nv.stream_ptr(dev)
        if plans.fwd.identity:
            dym = dout
        else:
            dym = torch.empty_like(dout)
            g_dev = g_hat.detach().float().contiguous()
            nv.check(lib.sd_node_mix_transposed(g_dev.data_ptr(), dout.data_ptr(), dym.data_ptr(), B, N, OUT, st), "sd_node_mix_transposed")
        dx = plans.bwd.forward(dym, precision=ctx.precision) if ctx.needs_input_grad[0] else None
        need_w, need_b, need_g = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        dW = db = dG = None
        if need_w or need_b or need_g:
            n_types = plans.fwd.n_types
            dW3 = torch.empty(n_types, OUT, plans.fwd.in_features, device=dev) if need_w else None
            db2 = torch.empty(n_types, OUT, device=dev) if need_b else None
            dG = torch.empty(N, N, device=dev) if need_g else None
            y_raw = plans.raw.forward(x, precision=ctx.precision) if need_g else None
            b2 = None
            if need_g and ctx.has_bias:
                b2 = bias.detach().float()
                b2 = (b2.unsqueeze(0) if b2.dim() == 1 else b2).contiguous()
            scratch = Workspace.get(dev, lib.sd_glin_backward_scratch_bytes(plans.fwd.handle, B), "glin_bwd")
            nv.check(lib.sd_glin_backward_params(plans.fwd.handle, x.data_ptr(), dym.data_ptr(), dout.data_ptr(), nv.dptr(y_raw), nv.dptr(b2),
                                                 nv.dptr(dW3), nv.dptr(db2), nv.dptr(dG), scratch.data_ptr(), B, 0, nv.stream_ptr(dev)),
                     "sd_glin_backward_params")
            dW = None if dW3 is None else dW3.reshape(weight.shape)
            db = None if db2 is None else db2.reshape(bias.shape)
        return dx, dW, db, dG, None, None


def graph_linear(layer, x: torch.Tensor, precision: str = "fp32") -> torch.Tensor:
    g = layer.G
    g_hat = F.normalize(g, p=1.0, dim=1) if layer.learn_influence else g        # graph_structural.py:31-34 (torch: N x N)
    return _GraphLinearFn.apply(x, layer.weight, layer.bias, g_hat, layer, _gemm_precision(precision))


class _SsTanhFn(torch.autograd.Function):
    """h = tanh(y (scale[t] + 1) + shift[t])   (attention.py:70-75); ss_table [T, 2C] or None"""

    @staticmethod
    def forward(ctx, y, ss_table, t32):
        y = y.contiguous()
        B, N, Cw = y.shape
        h = torch.empty_like(y)
        tab = None if ss_table is None else ss_table.detach().float().contiguous()
        nv.check(nv.load().sd_ss_tanh_forward(y.data_ptr(), nv.dptr(tab), nv.dptr(t32), h.data_ptr(), B, N, Cw, nv.stream_ptr(y.device)),
                 "sd_ss_tanh_forward")
        ctx.save_for_backward(y, h, tab if tab is not None else y.new_empty(0), t32 if t32 is not None else y.new_empty(0))
        ctx.has_ss = tab is not None
        return h

    @staticmethod
    def backward(ctx, dh):
        y, h, tab, t32 = ctx.saved_tensors
        dh = dh.contiguous()
        B, N, Cw = y.shape
        dy = torch.empty_like(y)
        rows = torch.empty(B, 2 * Cw, device=y.device) if ctx.has_ss else None
        nv.check(nv.load().sd_ss_tanh_backward(dh.data_ptr(), h.data_ptr(), y.data_ptr(), nv.dptr(tab) if ctx.has_ss else None,
                                               nv.dptr(t32) if ctx.has_ss else None, dy.data_ptr(), nv.dptr(rows), B, N, Cw,
                                               nv.stream_ptr(y.device)), "sd_ss_tanh_backward")
        dtab = None
        if ctx.has_ss and ctx.needs_input_grad[1]:
            # per-step sums of the per-sample rows: a [B, 2C] -> [T, 2C] segment sum, in sample order (deterministic)
            order = torch.argsort(t32.long(), stable=True)
            dtab = torch.zeros_like(tab)
            counts = torch.bincount(t32.long(), minlength=tab.shape[0])
            seg = torch.segment_reduce(rows[order], "sum", lengths=counts, unsafe=True)
            dtab.copy_(seg)
        return dy, dtab, None


class _RMSNormFn(torch.autograd.Function):
    """y = x / max(|x|, 1e-12) * g * sqrt(C)   (attention.py:30-36)"""

    @staticmethod
    def forward(ctx, x, g):
        x = x.contiguous()
        rows, Cw = x.numel() // x.shape[-1], x.shape[-1]
        y, inv = torch.empty_like(x), torch.empty(rows, device=x.device)
        gv = g.detach().float().reshape(-1).contiguous()
        nv.check(nv.load().sd_rmsnorm_forward(x.data_ptr(), gv.data_ptr(), y.data_ptr(), inv.data_ptr(), rows, Cw, nv.stream_ptr(x.device)),
                 "sd_rmsnorm_forward")
        ctx.save_for_backward(x, inv, gv)
        ctx.g_shape = g.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, inv, gv = ctx.saved_tensors
        dy = dy.contiguous()
        rows, Cw = inv.numel(), x.shape[-1]
        lib = nv.load()
        dx = torch.empty_like(x)
        part = torch.empty(lib.sd_rmsnorm_backward_blocks(rows), Cw, device=x.device)
        nv.check(lib.sd_rmsnorm_backward(dy.data_ptr(), x.data_ptr(), inv.data_ptr(), gv.data_ptr(), dx.data_ptr(), part.data_ptr(), rows, Cw,
                                         nv.stream_ptr(x.device)), "sd_rmsnorm_backward")
        return dx, part.sum(0).reshape(ctx.g_shape)


class _NodeAttentionFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(dh)) v over the nodes of a sample, per head   (attention.py:121-136)"""

    @staticmethod
    def forward(ctx, qkv, heads, dim_head):
        qkv = qkv.contiguous()
        B, N, _ = qkv.shape
        out = torch.empty(B, N, heads * dim_head, device=qkv.device)
        nv.check(nv.load().sd_node_attention(qkv.data_ptr(), out.data_ptr(), B, N, heads, dim_head, nv.stream_ptr(qkv.device)), "sd_node_attention")
        ctx.save_for_backward(qkv)
        ctx.hd = (heads, dim_head)
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        dout = dout.contiguous()
        B, N, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        nv.check(nv.load().sd_node_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, N, ctx.hd[0], ctx.hd[1],
                                                      nv.stream_ptr(qkv.device)), "sd_node_attention_backward")
        return dqkv, None, None


class _MahalanobisL1Fn(torch.autograd.Function):
    """loss_b = mean |S[t_b] (out_b - x0_b)|   (nonisotropic.py:176-190, base.py:297-298)"""

    @staticmethod
    def forward(ctx, out, x0, t32, s_tab):
        out, x0 = out.contiguous(), x0.contiguous()
        B, N, D = out.shape
        loss = torch.empty(B, device=out.device)
        nv.check(nv.load().sd_mahalanobis_loss(out.data_ptr(), x0.data_ptr(), t32.data_ptr(), s_tab.data_ptr(), loss.data_ptr(), B, N, D,
                                               nv.stream_ptr(out.device)), "sd_mahalanobis_loss")
        ctx.save_for_backward(out, x0, t32, s_tab)
        return loss

    @staticmethod
    def backward(ctx, gl):
        out, x0, t32, s_tab = ctx.saved_tensors
        B, N, D = out.shape
        dout = torch.empty_like(out)
        gl = gl.contiguous().float()
        nv.check(nv.load().sd_mahalanobis_loss_backward(out.data_ptr(), x0.data_ptr(), t32.data_ptr(), s_tab.data_ptr(), gl.data_ptr(),
                                                        dout.data_ptr(), B, N, D, nv.stream_ptr(out.device)), "sd_mahalanobis_loss_backward")
        return dout, None, None, None


# ---------------------------------------------------------------------------------------------------- Denoiser with autograd
def _sinusoidal(steps: torch.Tensor, dim: int, theta: float) -> torch.Tensor:
    import math
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=steps.device, dtype=torch.float32) * -(math.log(theta) / (half - 1)))
    ang = steps.float()[:, None] * freq[None]
    return torch.cat([ang.sin(), ang.cos()], -1)


def _resnet_block(blk, x, temb_table, t32, precision):
    ss = None
    if blk.mlp is not None and temb_table is not None:
        ss = blk.mlp(temb_table)                                   # [T, 2C]: tanh -> Linear on the T distinct steps (attention.py:81-84,93-96)
    h = _SsTanhFn.apply(graph_linear(blk.block1.proj, x, precision), ss, t32)
    h = _SsTanhFn.apply(graph_linear(blk.block2.proj, h, precision), None, None)
    res = x if isinstance(blk.res_linear, nn.Identity) else graph_linear(blk.res_linear, x, precision)
    return h + res


def _attention(att, x, precision):
    """Residual(PreNorm(Attention)) (attention.py:11-17,38-46,121-136)"""
    pre, attn = att.fn, att.fn.fn
    y = _RMSNormFn.apply(x, pre.norm.g)
    qkv = graph_linear(attn.to_qkv, y, precision)
    o = _NodeAttentionFn.apply(qkv, attn.heads, attn.dim_head)
    return graph_linear(attn.to_out, o, precision) + x


def denoiser_forward_train(model, x: torch.Tensor, t: torch.Tensor, x_cond: Optional[torch.Tensor] = None, num_steps: Optional[int] = None,
                           precision: str = "fp32") -> torch.Tensor:
    """Denoiser.forward (nn/generator.py:86-107) as a differentiable composition of the library's kernels.
    t: integer diffusion steps [B]; num_steps: rows of the time table (default max(t) + 1)."""
    nv.require_cuda(x, "x")
    t32 = t.to(x.device, torch.int32).contiguous()
    T = int(num_steps) if num_steps is not None else int(t32.max().item()) + 1
    if x_cond is not None and x_cond.shape[0] != x.shape[0]:                     # conditioning shared by consecutive rows (samples of a window)
        assert x.shape[0] % x_cond.shape[0] == 0, (x.shape, x_cond.shape)
        x_cond = x_cond.repeat_interleave(x.shape[0] // x_cond.shape[0], dim=0)
    xin = torch.cat([x_cond, x], -1) if x_cond is not None else x               # generator.py:91-92
    h = graph_linear(model.init_lin, xin.float().contiguous(), precision)
    r = h
    c = model.dim + model.cond_dim
    steps = torch.arange(T, device=x.device)
    temb = model.time_mlp[3](F.gelu(model.time_mlp[1](_sinusoidal(steps, c, model.theta))))     # [T, time_dim]  (generator.py:47-55)
    for blk, att in model.layers:                                                # generator.py:100-102
        h = _resnet_block(blk, h, temb, t32, precision)
        if not isinstance(att, nn.Identity):
            h = _attention(att, h, precision)
    h = torch.cat([h, r], -1)                                                    # :104
    h = _resnet_block(model.final_res_block, h, temb, t32, precision)            # :106
    return graph_linear(model.final_glin, h, precision)                          # :107


def diffusion_loss_train(diffusion, x_noisy: torch.Tensor, x_start: torch.Tensor, t: torch.Tensor, x_cond: Optional[torch.Tensor]):
    """(loss [B], model_out) with autograd through the Denoiser for the given (already noised) rows."""
    out = denoiser_forward_train(diffusion.model, x_noisy, t, x_cond, num_steps=diffusion.num_timesteps, precision=diffusion.precision)
    t32 = t.to(x_noisy.device, torch.int32).contiguous()
    loss = _MahalanobisL1Fn.apply(out, x_start.float().contiguous(), t32, diffusion.mahalanobis_S_sqrt_recip.contiguous())
    return loss, out


class SparseRowLoss(torch.autograd.Function):
    """loss [R] whose values were computed by the inference kernels; backward differentiates only the rows with a non-zero
    incoming gradient (see the module docstring).  `params` are the Denoiser's parameters: they are inputs of this node so that
    autograd routes their gradients here."""

    @staticmethod
    def forward(ctx, loss_values, diffusion, x_noisy, x_start, t, x_cond, rep, *params):
        ctx.diffusion, ctx.rep = diffusion, rep
        ctx.save_for_backward(x_noisy, x_start, t, x_cond if x_cond is not None else x_noisy.new_empty(0))
        ctx.has_cond = x_cond is not None
        ctx.n_params = len(params)
        return loss_values.clone()

    @staticmethod
    def backward(ctx, gl):
        x_noisy, x_start, t, x_cond = ctx.saved_tensors
        diffusion = ctx.diffusion
        rows = torch.nonzero(gl, as_tuple=False).flatten()
        params = [p for p in diffusion.model.parameters()]
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        if rows.numel():
            cond = None
            if ctx.has_cond:
                cond = x_cond[torch.div(rows, ctx.rep, rounding_mode="floor")] if ctx.rep > 1 else x_cond[rows]
            with torch.enable_grad():
                loss, _ = diffusion_loss_train(diffusion, x_noisy[rows].contiguous(), x_start[rows].contiguous(), t[rows].contiguous(), cond)
                live = [p for p in params if p.requires_grad]
                got = torch.autograd.grad(loss, live, gl[rows], allow_unused=True)
            it = iter(got)
            grads = [next(it) if p.requires_grad else None for p in params]
        return (None, None, None, None, None, None, None, *grads)


def ksimilarity_loss(diffusion_loss: torch.Tensor, batch: int, similarity: Optional[torch.Tensor] = None):
    """TrainerDiffusion.get_ksimilarity_loss (trainer.py:205-221): per observation the loss of the sample that is closest to the
    ground truth under `similarity` [batch * k] (default: the diffusion loss itself, similarity_space = 'latent_space')."""
    with torch.no_grad():
        sim = diffusion_loss if similarity is None else similarity
        idx = sim.view(batch, -1).min(dim=-1).indices
    return torch.gather(diffusion_loss.view(batch, -1), 1, idx.unsqueeze(1)).squeeze(-1), idx
