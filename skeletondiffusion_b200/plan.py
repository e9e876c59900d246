"""Host-side packing of module parameters into device buffers + native handles ("plans").

Packing is one-off plumbing per set of weights (PyTorch is used for the tiny [N,N] / [N,out]
algebra); the per-call work is all in the CUDA library.  Plans are rebuilt automatically when a
parameter changes (load_state_dict, .to(device), in-place updates) via `params_key`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _native as nv

__all__ = ["params_key", "invalidate_plans", "GlinPlan", "DenoiserPlan", "GruPlan", "Workspace"]


_EPOCH = 0        # bumped by invalidate_plans(): part of every key


def params_key(tensors: Sequence[Optional[torch.Tensor]]):
    """Identity of a set of parameters: storage address, autograd version counter and device of each, plus the global epoch.
    In-place writes through `.data` (`w.data.uniform_()`, `weight.data[1:] = ...`) do NOT bump the version counter: call
    `invalidate_plans()` after such a write (optimizer steps, load_state_dict, .to() and `with torch.no_grad(): p.copy_()` do)."""
    return (_EPOCH,) + tuple((t.data_ptr(), t._version, str(t.device)) if t is not None else None for t in tensors)


def invalidate_plans() -> None:
    """Force every packed plan (weight copies, operand planes, time tables, captured graphs keyed on them) to be rebuilt at its next
    use.  Needed only after parameter writes that bypass autograd's version counter (`.data` assignments)."""
    global _EPOCH
    _EPOCH += 1


def _normalized_influence(G: torch.Tensor, learn_influence: bool) -> torch.Tensor:
    # graph_structural.py:31-35: row-L1 normalisation only when the influence matrix is learnable
    return F.normalize(G, p=1.0, dim=1) if learn_influence else G


def _f16x2_planes(w: torch.Tensor) -> torch.Tensor:
    """[2, ...] fp16 planes of the two-plane operand split (SD_PREC_F16X2): w ~= hi + lo * 2^-11 with hi = fp16(w),
    lo = fp16((w - hi) * 2^11); the residual is exact in fp32, so the pair carries 22 significand bits."""
    hi = w.to(torch.float16)
    lo = ((w - hi.float()) * 2048.0).to(torch.float16)
    return torch.stack([hi, lo], 0).contiguous()


def gate_interleave_perm(hidden: int, device=None) -> torch.Tensor:
    """Row order of the fused GRU-step kernels: new row 96 blk + 32 g + u holds original row g H + 32 blk + u (g = gate r / z / n of
    `weight_hh` / `weight_ih`, recurrent.py:333-349), so every 96-row block - one n-tile of the GEMM - carries the three gates of
    hidden units [32 blk, 32 blk + 32)."""
    if hidden % 32:
        raise ValueError("gate-interleaved GRU weights need a hidden size that is a multiple of 32")
    blk = torch.arange(hidden // 32, device=device).view(-1, 1, 1)
    g = torch.arange(3, device=device).view(1, -1, 1)
    u = torch.arange(32, device=device).view(1, 1, -1)
    return (g * hidden + 32 * blk + u).reshape(-1)


def _is_identity(g: torch.Tensor) -> bool:
    return bool(torch.equal(g, torch.eye(g.shape[0], device=g.device, dtype=g.dtype)))


class Workspace:
    """Grow-only per-device scratch buffer handed to the composite kernels (caller-owned memory)."""
    _bufs: Dict[str, torch.Tensor] = {}

    @classmethod
    def get(cls, device: torch.device, nbytes: int, tag: str = "main") -> torch.Tensor:
        key = f"{device}:{tag}"
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf


class GlinPlan:
    """Packed StaticGraphLinear: fp32 weights [types,out,in], bias_node = G^ @ bias[type], G^ (or identity)."""

    def __init__(self, num_nodes: int, node_types: Optional[torch.Tensor], weight: torch.Tensor,
                 bias: Optional[torch.Tensor], g_hat: torch.Tensor, key=None):
        nv.require_cuda(weight, "weight")
        dev = weight.device
        self.key, self.device, self.N = key, dev, num_nodes
        w = weight.detach().to(torch.float32)
        if w.dim() == 2:
            w = w.unsqueeze(0)
        self.weight = w.contiguous()
        self.n_types, self.out_features, self.in_features = self.weight.shape
        types = node_types.to("cpu", torch.int64).tolist() if node_types is not None else [0] * num_nodes
        self.types_host = (C.c_int32 * num_nodes)(*types)
        g_hat = g_hat.detach().to(dev, torch.float32)
        self.identity = _is_identity(g_hat)
        self.g = None if self.identity else g_hat.contiguous()
        self.bias_node = None
        if bias is not None:
            b = bias.detach().to(torch.float32)
            b_node = b[torch.tensor(types, device=dev)] if b.dim() == 2 else b.unsqueeze(0).expand(num_nodes, -1)
            self.bias_node = (g_hat @ b_node).contiguous()      # bias is added before the mix (graph_structural.py:38-41)
        self.handle = C.c_void_p()
        nv.check(nv.load().sd_glin_create(num_nodes, self.types_host, self.n_types, self.in_features, self.out_features,
                                          self.weight.data_ptr(), nv.dptr(self.bias_node), nv.dptr(self.g),
                                          C.byref(self.handle)), "sd_glin_create")
        # K-major fp32 copy for the FFMA2 kernel (an output pair is then one 8-byte shared-memory load)
        self.weight_kmajor = None
        if self.out_features % 96 == 0 and self.in_features % 32 == 0:
            self.weight_kmajor = self.weight.transpose(1, 2).contiguous()
            nv.check(nv.load().sd_glin_set_kmajor(self.handle, self.weight_kmajor.data_ptr()), "sd_glin_set_kmajor")
        # bf16 weight planes for the tcgen05 paths (K-major [3, types, out, in], read by TMA): w = p0 + p1 + p2 exactly;
        # plane 0 alone (= bf16(w)) is what the plain bf16 path multiplies with
        p0 = self.weight.to(torch.bfloat16)
        r1 = self.weight - p0.float()
        p1 = r1.to(torch.bfloat16)
        p2 = (r1 - p1.float()).to(torch.bfloat16)
        self.weight_bf16 = torch.stack([p0, p1, p2], 0).contiguous()
        nv.check(nv.load().sd_glin_set_bf16(self.handle, self.weight_bf16.data_ptr(), 3), "sd_glin_set_bf16")
        self.weight_f16 = _f16x2_planes(self.weight)
        nv.check(nv.load().sd_glin_set_f16x2(self.handle, self.weight_f16.data_ptr()), "sd_glin_set_f16x2")

    @classmethod
    def from_layer(cls, layer, fold_in: Optional[torch.Tensor] = None, key=None) -> "GlinPlan":
        w = layer.weight.detach()
        if fold_in is not None:          # RMSNorm gain folded into the input columns of the weight
            w = w * fold_in.detach().to(w.device, w.dtype)
        g_hat = _normalized_influence(layer.G.detach(), layer.learn_influence)
        return cls(layer.num_nodes, layer.node_type_index, w, None if layer.bias is None else layer.bias.detach(), g_hat, key=key)

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                nv.load().sd_glin_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def forward(self, x: torch.Tensor, x2: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
                scale_shift: Optional[torch.Tensor] = None, ss_rows: Optional[torch.Tensor] = None, act: int = nv.ACT_NONE,
                residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, rep: int = 1,
                precision: str = "fp32") -> torch.Tensor:
        nv.require_cuda(x, "input")
        if x.dtype != torch.float32:
            x = x.float()
        if x.dim() != 3:
            raise ValueError(f"expected [B, N, C] input, got {tuple(x.shape)}")
        if x.stride(-1) != 1:
            x = x.contiguous()
        batch = x.shape[0] * rep
        if out is None:
            out = torch.empty(batch, self.N, self.out_features, device=x.device, dtype=torch.float32)
        args = nv.SdGlinArgs()
        args.a0 = nv.view_of(x, rep)
        args.a1 = nv.view_of(x2 if x2 is None or x2.stride(-1) == 1 else x2.contiguous())
        args.row_scale_dev = nv.dptr(row_scale)
        args.scale_shift_dev = nv.dptr(scale_shift)
        args.ss_row_dev = nv.dptr(ss_rows)
        args.ss_row = 0
        args.ss_row_stride = 0 if scale_shift is None else scale_shift.stride(0)
        args.act = act
        args.residual = nv.view_of(residual)
        args.out = nv.view_of(out)
        scratch = None
        if precision == "bf16":       # bf16 operand copy + fp32 raw product (see glin_forward_tc)
            scratch = Workspace.get(x.device, batch * self.N * (2 * self.in_features + 4 * self.out_features) + 1024, "glin")
        elif not self.identity or (precision in nv.FP32_GRADE_TC and x2 is not None):
            # pre-mix product (non-identity G^) or the partial product of a K-split two-segment layer (bf16x3)
            scratch = Workspace.get(x.device, batch * self.N * self.out_features * 4, "glin")
        args.scratch_dev = nv.dptr(scratch)
        args.batch = batch
        args.precision = nv.PRECISIONS[precision]
        nv.check(nv.load().sd_glin_forward(self.handle, C.byref(args), nv.stream_ptr(x.device)), "sd_glin_forward")
        return out


class DenoiserPlan:
    """All 8*depth+5 graph-linears of a Denoiser + the batch-invariant time-conditioning table."""

    def __init__(self, model, key=None, time_rows: int = 16):
        lib = nv.load()
        p0 = next(model.parameters())
        nv.require_cuda(p0, "Denoiser parameters")
        dev = p0.device
        self.key, self.device, self.time_rows = key, dev, time_rows
        self.N, self.C, self.out_dim = model.channels, model.dim + model.cond_dim, model.out_dim
        self.dim, self.cond_dim = model.dim, model.cond_dim
        self.handle = C.c_void_p()
        nv.check(lib.sd_denoiser_create(self.N, model.dim, model.cond_dim, model.out_dim, model.depth, model.heads,
                                        model.dim_head, C.byref(self.handle)), "sd_denoiser_create")
        self.layers: List[GlinPlan] = []

        def put(slot: int, layer, fold_in=None):
            plan = GlinPlan.from_layer(layer, fold_in=fold_in)
            self.layers.append(plan)
            nv.check(lib.sd_denoiser_set_layer(self.handle, slot, plan.handle), "sd_denoiser_set_layer")

        put(0, model.init_lin)
        heads_w, heads_b = [], []
        n_pairs = 2 * model.depth
        for i, (blk, att) in enumerate(model.layers):
            put(1 + 4 * i, blk.block1.proj)
            put(2 + 4 * i, blk.block2.proj)
            heads_w.append(blk.mlp[1].weight)
            heads_b.append(blk.mlp[1].bias)
            if i != n_pairs - 1:
                pre = att.fn                       # Residual.fn = PreNorm
                fold = pre.norm.g.detach().reshape(-1) * (self.C ** 0.5)
                put(3 + 4 * i, pre.fn.to_qkv, fold_in=fold)
                put(4 + 4 * i, pre.fn.to_out)
        sf = 1 + 8 * model.depth
        fin = model.final_res_block
        put(sf, fin.block1.proj)
        put(sf + 1, fin.block2.proj)
        put(sf + 2, fin.res_linear)
        put(sf + 3, model.final_glin)
        heads_w.append(fin.mlp[1].weight)
        heads_b.append(fin.mlp[1].bias)
        # time table for integer times 0..time_rows-1  (generator.py:47-55, attention.py:81-84)
        n_heads = len(heads_w)
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self._keep = [f32(t) for t in (model.time_mlp[1].weight, model.time_mlp[1].bias, model.time_mlp[3].weight,
                                        model.time_mlp[3].bias)] + [f32(t) for t in heads_w] + [f32(t) for t in heads_b]
        hw = (C.c_void_p * n_heads)(*[t.data_ptr() for t in self._keep[4:4 + n_heads]])
        hb = (C.c_void_p * n_heads)(*[t.data_ptr() for t in self._keep[4 + n_heads:]])
        times = torch.arange(time_rows, device=dev, dtype=torch.float32)
        self.table = torch.empty(time_rows, n_heads, 2 * self.C, device=dev, dtype=torch.float32)
        ws = torch.empty(time_rows * (self.C + 2 * model.time_dim), device=dev, dtype=torch.float32)
        nv.check(lib.sd_time_table(times.data_ptr(), time_rows, self.C, model.theta, model.time_dim,
                                   self._keep[0].data_ptr(), self._keep[1].data_ptr(), self._keep[2].data_ptr(),
                                   self._keep[3].data_ptr(), hw, hb, n_heads, self.table.data_ptr(), ws.data_ptr(),
                                   nv.stream_ptr(dev)), "sd_time_table")
        torch.cuda.current_stream(dev).synchronize()      # ws / times may be freed after this point
        nv.check(lib.sd_denoiser_set_time_table(self.handle, self.table.data_ptr(), time_rows), "sd_denoiser_set_time_table")

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                nv.load().sd_denoiser_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def workspace(self, batch: int, precision: str) -> torch.Tensor:
        n = nv.load().sd_denoiser_workspace_bytes(self.handle, batch, nv.PRECISIONS[precision])
        return Workspace.get(self.device, n, "denoiser")

    def forward(self, x: torch.Tensor, x_cond: Optional[torch.Tensor], t_rows: torch.Tensor, precision: str = "fp32",
                cond_rep: Optional[int] = None) -> torch.Tensor:
        x = x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        B = x.shape[0]
        if tuple(x.shape[1:]) != (self.N, self.dim):
            raise ValueError(f"Denoiser expects [B, {self.N}, {self.dim}], got {tuple(x.shape)}")
        rep = 1
        if x_cond is not None:
            x_cond = x_cond.float()
            if x_cond.stride(-1) != 1:
                x_cond = x_cond.contiguous()
            if x_cond.shape[0] != B:           # fewer conditioning rows: repeat_interleave in place (base.py:246-248)
                if B % x_cond.shape[0] != 0:
                    raise ValueError("batch must be a multiple of the number of conditioning rows")
                rep = B // x_cond.shape[0]
        out = torch.empty(B, self.N, self.out_dim, device=x.device, dtype=torch.float32)
        ws = self.workspace(B, precision)
        xv, cv = nv.view_of(x), nv.view_of(x_cond, rep)
        if isinstance(t_rows, int):                     # every sample at the same diffusion time (the sampling loop's case)
            rows_ptr, row0 = None, t_rows
        else:
            t_rows = t_rows.to(x.device, torch.int32).contiguous()
            rows_ptr, row0 = t_rows.data_ptr(), 0
        nv.check(nv.load().sd_denoiser_forward(self.handle, C.byref(xv), C.byref(cv) if x_cond is not None else None,
                                               rows_ptr, row0, out.data_ptr(), B, ws.data_ptr(),
                                               nv.PRECISIONS[precision], nv.stream_ptr(x.device)), "sd_denoiser_forward")
        return out


class GruPlan:
    """Packed StaticGraphGRU cell for `steps` recurrent steps (recurrent.py:321-366).

    The graph-influence sequence gx_i is data independent (gx_0 = normalize(G),
    gx_{i+1} = normalize(gx_i + G_add), recurrent.py:325-327, 361-363), so it is tabulated here
    together with the mixed biases gx_i @ bias[type]."""

    def __init__(self, cell, steps: int, key=None):
        nv.require_cuda(cell.weight_ih, "GRU parameters")
        dev = cell.weight_ih.device
        self.key, self.steps, self.device = key, steps, dev
        N, H = cell.num_nodes, cell.hidden_size
        types = cell.node_type_index.to("cpu", torch.int64).tolist() if cell.node_type_index is not None else [0] * N
        self.types_host = (C.c_int32 * N)(*types)
        w_ih, w_hh = cell.weight_ih.detach().float(), cell.weight_hh.detach().float()
        if w_ih.dim() == 2:
            w_ih, w_hh = w_ih.unsqueeze(0), w_hh.unsqueeze(0)
        self.w_ih, self.w_hh = w_ih.contiguous(), w_hh.contiguous()
        tix = torch.tensor(types, device=dev)

        def node_bias(b):
            if b is None:
                return torch.zeros(N, 3 * H, device=dev)
            b = b.detach().float()
            return b[tix] if b.dim() == 2 else b.unsqueeze(0).expand(N, -1)

        b_ih, b_hh = node_bias(cell.bias_ih), node_bias(cell.bias_hh)
        G = cell.G.detach().float()
        g_add = cell.G_add.detach().float() if torch.is_tensor(cell.G_add) else float(cell.G_add)
        renorm = cell.learn_influence or cell.learn_additive_graph_influence
        gx = F.normalize(G, p=1.0, dim=1) if cell.learn_influence else G
        seq = []
        for _ in range(steps):
            seq.append(gx)
            gx = gx + g_add
            if renorm:
                gx = F.normalize(gx, p=1.0, dim=1)
        gx_seq = torch.stack(seq, 0).contiguous()                      # [steps, N, N]
        eye = torch.eye(N, device=dev)
        self.identity = bool((gx_seq == eye).all())
        self.gx_seq = None if self.identity else gx_seq
        self.bias_ih_seq = (gx_seq @ b_ih).contiguous()                # [steps, N, 3H]
        self.bias_hh_seq = (gx_seq @ b_hh).contiguous()
        self.handle = C.c_void_p()
        nv.check(nv.load().sd_gru_create(N, self.types_host, self.w_ih.shape[0], cell.input_size, H, self.w_ih.data_ptr(),
                                         self.w_hh.data_ptr(), self.bias_ih_seq.data_ptr(), self.bias_hh_seq.data_ptr(),
                                         nv.dptr(self.gx_seq), steps, C.byref(self.handle)), "sd_gru_create")
        # bf16 planes of W_hh for the tcgen05 recurrent product (w = p0 + p1 + p2 exactly), K-major [3, types, 3H, H]
        p0 = self.w_hh.to(torch.bfloat16)
        r1 = self.w_hh - p0.float()
        p1 = r1.to(torch.bfloat16)
        p2 = (r1 - p1.float()).to(torch.bfloat16)
        self.w_hh_planes = torch.stack([p0, p1, p2], 0).contiguous()
        nv.check(nv.load().sd_gru_set_bf16x3(self.handle, self.w_hh_planes.data_ptr()), "sd_gru_set_bf16x3")
        self.w_hh_f16 = _f16x2_planes(self.w_hh)
        nv.check(nv.load().sd_gru_set_f16x2(self.handle, self.w_hh_f16.data_ptr()), "sd_gru_set_f16x2")
        if self.identity and H % 32 == 0:
            # gate-interleaved copies for the fused FFMA2 GRU step: every 96-row block = gates r|z|n of 32 units
            perm = gate_interleave_perm(H, dev)                             # new row -> original row
            self.w_ih_perm = self.w_ih[:, perm].contiguous()
            self.w_hh_perm = self.w_hh[:, perm].transpose(1, 2).contiguous()     # K-major [types, H, 3H]
            self.bias_ih_perm = b_ih[:, perm].contiguous()
            self.bias_hh_perm = b_hh[:, perm].contiguous()
            nv.check(nv.load().sd_gru_set_fused(self.handle, self.w_ih_perm.data_ptr(), self.w_hh_perm.data_ptr(),
                                                self.bias_ih_perm.data_ptr(), self.bias_hh_perm.data_ptr()), "sd_gru_set_fused")
            # the same row order as fp16 planes [2, types, 3H, H]: fused tensor-core step (product + gates in one tcgen05 kernel)
            self.w_hh_perm_f16 = _f16x2_planes(self.w_hh[:, perm].contiguous())
            nv.check(nv.load().sd_gru_set_fused_f16x2(self.handle, self.w_hh_perm_f16.data_ptr()), "sd_gru_set_fused_f16x2")

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                nv.load().sd_gru_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass
