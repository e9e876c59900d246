"""Graph layers and the Denoiser with the reference's constructor signatures and state_dict keys.

The modules below only *hold parameters* (same names / shapes / nesting as the reference, so
`cvpr_release.pt`-style checkpoints load with strict=True); every forward runs hand-written sm_100a
kernels through the C ABI (skeletondiffusion_b200/_native.py).  Reference files:
  StaticGraphLinear  src/core/network/layers/graph_structural.py:58-114 (forward :30-43)
  RMSNorm/PreNorm/Residual/Block/ResnetBlock/Attention  src/core/network/layers/attention.py:11-136
  Denoiser           src/core/network/nn/generator.py:8-107
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _native as nv
from .plan import DenoiserPlan, GlinPlan, params_key

__all__ = ["StaticGraphLinear", "RMSNorm", "PreNorm", "Residual", "Block", "ResnetBlock", "Attention",
           "SinusoidalPosEmb", "Denoiser"]


class StaticGraphLinear(nn.Module):
    """out[b,n] = sum_m G^[n,m] (x[b,m] W[type(m)]^T + bias[type(m)])  (graph_structural.py:30-43)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, num_nodes: int = None,
                 graph_influence=None, learn_influence: bool = False, node_types: torch.Tensor = None,
                 weights_per_type: bool = False, **kwargs):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.learn_influence = learn_influence
        if graph_influence is not None:
            assert num_nodes is None or num_nodes == graph_influence.shape[0]
            num_nodes = graph_influence.shape[0]
            g0 = graph_influence
        else:
            assert num_nodes, "Number of Nodes or Graph Influence Matrix has to be given."
            g0 = torch.eye(num_nodes, num_nodes)
        if isinstance(g0, nn.Parameter):
            assert learn_influence
            self.G = g0
        elif learn_influence:
            self.G = nn.Parameter(g0)
        else:
            self.register_buffer("G", g0)
        self.num_nodes = num_nodes
        if weights_per_type and node_types is None:
            node_types = torch.arange(num_nodes)
        self.node_type_index = node_types          # plain attribute, not a buffer (graph_structural.py:100)
        if node_types is not None:
            n_types = int(node_types.max()) + 1
            self.weight = nn.Parameter(torch.empty(n_types, out_features, in_features))
            bias_shape = (n_types, out_features)
        else:
            self.weight = nn.Parameter(torch.empty(out_features, in_features))
            bias_shape = (out_features,)
        if bias:
            self.bias = nn.Parameter(torch.empty(*bias_shape))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._plan: Optional[GlinPlan] = None

    def reset_parameters(self) -> None:
        # same distribution as the reference (graph_structural.py:17-28): kaiming-uniform on the
        # whole tensor (fan_in = size(1) * receptive field), every node type starts identical
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.weight.dim() == 3:
            self.weight.data[1:] = self.weight.data[0]
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def plan(self, fold_gain: Optional[torch.Tensor] = None) -> GlinPlan:
        """fold_gain: RMSNorm gain `g` ([1,1,in]) folded (times sqrt(in)) into the weight's input columns."""
        key = params_key([self.G, self.weight, self.bias, fold_gain])
        if self._plan is None or self._plan.key != key:
            fold = None if fold_gain is None else fold_gain.detach().reshape(-1) * (self.in_features ** 0.5)
            self._plan = GlinPlan.from_layer(self, fold_in=fold, key=key)
        return self._plan

    def forward(self, input: torch.Tensor, g: Optional[torch.Tensor] = None, *, row_scale=None, act: int = nv.ACT_NONE,
                residual: Optional[torch.Tensor] = None, precision: str = "fp32") -> torch.Tensor:
        if g is not None:
            raise NotImplementedError("per-call graph influence `g` is not used on the sampling path")
        return self.plan().forward(input, row_scale=row_scale, act=act, residual=residual, precision=precision)


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *args, **kwargs):
        return self.fn(x, *args, residual=x, **kwargs)


class RMSNorm(nn.Module):
    """x / max(||x||, 1e-12) * g * sqrt(C) (attention.py:30-36).  On the kernel path g*sqrt(C) is folded
    into the following to_qkv weights and 1/||x|| is applied to the product rows."""

    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, 1, dim))

    def inv_norm(self, x: torch.Tensor) -> torch.Tensor:
        nv.require_cuda(x, "x")
        x = x.contiguous()
        inv = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
        nv.check(nv.load().sd_row_inv_norm(x.data_ptr(), inv.data_ptr(), inv.numel(), x.shape[-1], nv.stream_ptr(x.device)),
                 "sd_row_inv_norm")
        return inv


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = RMSNorm(dim)

    def forward(self, x, residual=None, **kwargs):
        return self.fn(x, norm=self.norm, residual=residual, **kwargs)


class Block(nn.Module):
    """tanh( SGL(x) * (scale + 1) + shift )  (attention.py:49-75); norm_type 'none', act 'tanh' only."""

    def __init__(self, dim, dim_out, norm_type="none", act_type="tanh", *args, **kwargs):
        super().__init__()
        kwargs.pop("groups", None)
        if norm_type != "none":
            raise NotImplementedError("norm_type='layer' is not on the shipped sampling path (configs use 'none')")
        if act_type != "tanh":
            raise NotImplementedError(f"act_type {act_type}")
        self.proj = StaticGraphLinear(dim, dim_out, *args, **kwargs)
        self.norm = nn.Identity()
        self.act = nn.Tanh()


class ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8, **kwargs):
        super().__init__()
        self.mlp = nn.Sequential(nn.Tanh(), nn.Linear(time_emb_dim, dim_out * 2)) if time_emb_dim is not None else None
        self.block1 = Block(dim, dim_out, groups=groups, **kwargs)
        self.block2 = Block(dim_out, dim_out, groups=groups, **kwargs)
        self.res_linear = StaticGraphLinear(dim, dim_out, bias=False, **kwargs) if dim != dim_out else nn.Identity()


class Attention(nn.Module):
    """Multi-head attention across the nodes of one sample (attention.py:105-136)."""

    def __init__(self, dim, dim_out=None, heads=4, dim_head=32, qkv_bias: bool = False, attn_dropout: float = 0.,
                 proj_dropout: float = 0., qk_norm: bool = False, norm_layer=nn.Identity, **kwargs):
        super().__init__()
        if qk_norm or attn_dropout or proj_dropout:
            raise NotImplementedError("qk_norm / dropout are not used by the shipped configs")
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        hidden = dim_head * heads
        self.to_qkv = StaticGraphLinear(dim, hidden * 3, bias=qkv_bias, **kwargs)
        self.to_out = StaticGraphLinear(hidden, dim_out if dim_out is not None else dim, bias=False, **kwargs)
        self.attn_dropout, self.out_dropout = nn.Dropout(attn_dropout), nn.Dropout(proj_dropout)
        self.q_norm, self.k_norm = nn.Identity(), nn.Identity()

    def forward(self, x: torch.Tensor, norm: Optional[RMSNorm] = None, residual: Optional[torch.Tensor] = None,
                precision: str = "fp32") -> torch.Tensor:
        nv.require_cuda(x, "x")
        x = x.contiguous()
        b, n, c = x.shape
        if norm is not None:
            qkv = self.to_qkv.plan(fold_gain=norm.g).forward(x, row_scale=norm.inv_norm(x), precision=precision)
        else:
            qkv = self.to_qkv(x, precision=precision)
        out = torch.empty(b, n, self.heads * self.dim_head, device=x.device, dtype=torch.float32)
        nv.check(nv.load().sd_node_attention(qkv.data_ptr(), out.data_ptr(), b, n, self.heads, self.dim_head,
                                             nv.stream_ptr(x.device)), "sd_node_attention")
        return self.to_out(out, residual=residual, precision=precision)


class SinusoidalPosEmb(nn.Module):
    """Parameter-free; computed inside sd_time_table.  Public definition of
    denoising-diffusion-pytorch==1.9.4 (imported by the reference at nn/generator.py:3)."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        self.dim, self.theta = dim, theta


class Denoiser(nn.Module):
    def __init__(self, dim, out_dim, channels: int, cond_dim: int = 0, depth=1, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000,
                 attn_dim_head=32, attn_heads=4, use_attention=True, **kwargs):
        super().__init__()
        if learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("learned/random sinusoidal conditioning is not used by any shipped config")
        if not use_attention:
            raise NotImplementedError("use_attention=False is not used by any shipped config")
        if self_condition:
            raise NotImplementedError("self_condition=True (broken in the reference's training path, base.py:280) is not supported")
        self.channels, self.self_condition = channels, self_condition
        self.dim, self.cond_dim, self.depth = dim, cond_dim, depth
        self.heads, self.dim_head = attn_heads, attn_dim_head
        c = dim + cond_dim
        self.init_lin = StaticGraphLinear(dim + cond_dim, c, bias=True, **kwargs)
        time_dim = c * 4
        self.time_dim, self.theta = time_dim, float(sinusoidal_pos_emb_theta)
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(c, theta=sinusoidal_pos_emb_theta), nn.Linear(c, time_dim),
                                      nn.GELU(), nn.Linear(time_dim, time_dim))
        self.layers = nn.ModuleList([])
        for i in range(depth):                                                  # generator.py:59-77
            for last in (False, i == depth - 1):
                blk = ResnetBlock(c, c, time_emb_dim=time_dim, groups=resnet_block_groups, **kwargs)
                att = nn.Identity() if last else Residual(PreNorm(c, Attention(c, heads=attn_heads, dim_head=attn_dim_head, **kwargs)))
                self.layers.append(nn.ModuleList([blk, att]))
        self.out_dim = out_dim * (1 if not learned_variance else 2)
        self.final_res_block = ResnetBlock(c * 2, c, time_emb_dim=time_dim, groups=resnet_block_groups, **kwargs)
        self.final_glin = StaticGraphLinear(c, self.out_dim, bias=True, **kwargs)
        self._plan: Optional[DenoiserPlan] = None

    # ------------------------------------------------------------------ plan (packed weights)
    def plan(self, min_time_rows: int = 1) -> DenoiserPlan:
        key = params_key(list(self.parameters()) + list(self.buffers()))
        if self._plan is None or self._plan.key != key or self._plan.time_rows < min_time_rows:
            rows = max(min_time_rows, self._plan.time_rows if self._plan is not None and self._plan.key == key else 0, 16)
            self._plan = DenoiserPlan(self, key=key, time_rows=rows)
        return self._plan

    def forward(self, x, time, x_self_cond=None, x_cond=None, precision: str = "fp32"):
        nv.require_cuda(x, "x")
        if x_cond is None and self.cond_dim > 0:
            raise ValueError("x_cond is required when cond_dim > 0")
        t = time.reshape(-1)
        if t.numel() != x.shape[0]:
            raise ValueError("time must have one entry per sample")
        if t.is_floating_point():
            if not bool((t == t.round()).all()):
                raise NotImplementedError("non-integer diffusion times")
        t_int = t.to(torch.int64)
        t_max = int(t_int.max().item()) if t.numel() else 0
        t_min = int(t_int.min().item()) if t.numel() else 0
        if t_min < 0:
            raise ValueError("negative diffusion time")
        plan = self.plan(min_time_rows=t_max + 1)
        # uniform time (every sampling step): one table row for the whole batch, no per-sample gather
        rows = t_max if t_min == t_max else t_int.to(torch.int32)
        return plan.forward(x, x_cond, rows, precision=precision)
