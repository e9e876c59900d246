"""Recurrent graph autoencoder with the reference's constructor signatures and state_dict keys.

Reference: AutoEncoder src/core/network/nn/autoencoder.py:8-98, Encoder nn/encoder.py:10-82,
Decoder nn/decoder.py:9-104, StaticGraphGRU(_Cell_) layers/recurrent.py:208-402.
Modules hold parameters only; encode/decode run as CUDA kernels through sd_encode / sd_decode.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from . import _native as nv
from .network import StaticGraphLinear
from .plan import GruPlan, Workspace, params_key

__all__ = ["StaticGraphGRUCell", "StaticGraphGRU", "Encoder", "Decoder", "AutoEncoder"]


class StaticGraphGRUCell(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_nodes: int = None, dropout: float = 0.,
                 recurrent_dropout: float = 0., graph_influence=None, learn_influence: bool = False,
                 additive_graph_influence=None, learn_additive_graph_influence: bool = False,
                 node_types: torch.Tensor = None, weights_per_type: bool = False, clockwork: bool = False,
                 bias: bool = True):
        super().__init__()
        if clockwork:
            raise NotImplementedError("clockwork=True is not used by the shipped configs")
        if dropout or recurrent_dropout:
            raise NotImplementedError("dropout is inactive on the sampling path")
        self.input_size, self.hidden_size = input_size, hidden_size
        self.learn_influence = learn_influence
        self.learn_additive_graph_influence = learn_additive_graph_influence
        if graph_influence is not None:
            num_nodes = graph_influence.shape[0]
            g0 = graph_influence
        else:
            assert num_nodes, "Number of Nodes or Graph Influence Matrix has to be given."
            g0 = torch.eye(num_nodes, num_nodes)
        if isinstance(g0, nn.Parameter) or learn_influence:
            self.G = g0 if isinstance(g0, nn.Parameter) else nn.Parameter(g0)
        else:
            self.register_buffer("G", g0)
        if additive_graph_influence is not None:
            if isinstance(additive_graph_influence, nn.Parameter) or learn_additive_graph_influence:
                self.G_add = additive_graph_influence if isinstance(additive_graph_influence, nn.Parameter) else nn.Parameter(additive_graph_influence)
            else:
                self.register_buffer("G_add", additive_graph_influence)
        elif learn_additive_graph_influence:
            self.G_add = nn.Parameter(torch.zeros_like(self.G))
        else:
            self.G_add = 0.
        if weights_per_type and node_types is None:
            node_types = torch.arange(num_nodes)
        shape = lambda *s: (int(node_types.max()) + 1, *s) if node_types is not None else s
        self.weight_ih = nn.Parameter(torch.empty(*shape(3 * hidden_size, input_size)))
        self.weight_hh = nn.Parameter(torch.empty(*shape(3 * hidden_size, hidden_size)))
        self.register_buffer("node_type_index", node_types)          # a buffer here (recurrent.py:276), unlike the linear layer
        if bias:
            self.bias_ih = nn.Parameter(torch.empty(*shape(3 * hidden_size)))
            self.bias_hh = nn.Parameter(torch.empty(*shape(3 * hidden_size)))
        else:
            self.bias_ih = self.bias_hh = None
        self.register_buffer("phase", torch.ones(hidden_size))       # clockwork=False (recurrent.py:304-306)
        self.num_nodes = num_nodes
        stdv = 1.0 / math.sqrt(hidden_size)                            # recurrent.py:312-319
        for name, w in self.named_parameters():
            if name not in ("G", "G_add"):
                w.data.uniform_(-stdv, stdv)
        self._plan: Optional[GruPlan] = None

    def plan(self, steps: int) -> GruPlan:
        key = params_key([self.G, self.G_add if torch.is_tensor(self.G_add) else None, self.weight_ih, self.weight_hh,
                          self.bias_ih, self.bias_hh])
        if self._plan is None or self._plan.key != key or self._plan.steps < steps:
            self._plan = GruPlan(self, steps=max(steps, self._plan.steps if self._plan is not None else 0), key=key)
        return self._plan


class StaticGraphGRU(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, layer_dropout: float = 0.0, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([StaticGraphGRUCell(input_size, hidden_size, **kwargs)] +
                                    [StaticGraphGRUCell(hidden_size, hidden_size, **kwargs) for _ in range(num_layers - 1)])
        self.dropout = nn.Dropout(layer_dropout)


class Encoder(nn.Module):
    def __init__(self, num_nodes: int, input_size: int, hidden_size: int, output_size: int, node_types: torch.Tensor = None,
                 enc_num_layers: int = 1, dropout: float = 0., encoder_act: str = "tanh", recurrent_arch: str = "StaticGraphGRU", **kwargs):
        super().__init__()
        if recurrent_arch != "StaticGraphGRU":
            raise NotImplementedError("only StaticGraphGRU is on the shipped path (LSTM variant is out of scope)")
        assert encoder_act in ("tanh", "identity")
        self.encoder_act = encoder_act
        self.activation_fn = nn.Tanh() if encoder_act == "tanh" else nn.Identity()
        self.num_layers, self.recurrent_arch = enc_num_layers, recurrent_arch
        self.rnn = StaticGraphGRU(input_size, hidden_size, num_layers=enc_num_layers, node_types=node_types,
                                  num_nodes=num_nodes, bias=True, clockwork=False, learn_influence=True)
        self.fc = StaticGraphLinear(hidden_size, output_size, num_nodes=num_nodes, node_types=node_types, bias=True, learn_influence=True)
        self.initial_hidden1 = StaticGraphLinear(input_size, hidden_size, num_nodes=num_nodes, node_types=node_types, bias=True, learn_influence=True)
        self.dropout = nn.Dropout(dropout)

    def encode(self, x: torch.Tensor, final_act: int, precision: str = "fp32") -> torch.Tensor:
        """[W, T, N, F] -> [W, N, latent]; final_act is applied on top of fc (ACT_* code).  precision 'bf16x3' / 'bf16' runs
        the recurrent products on the tcgen05 3-plane kernel (fp32-grade), 'fp32' on the FFMA kernels."""
        nv.require_cuda(x, "observation")
        x = x.float().contiguous()
        W, T, N, Fdim = x.shape
        cells = list(self.rnn.layers)
        H = cells[0].hidden_size
        plans = [c.plan(T) for c in cells]
        handles = (C.c_void_p * len(plans))(*[p.handle.value for p in plans])
        ih, fc = self.initial_hidden1.plan(), self.fc.plan()
        lib = nv.load()
        ws = Workspace.get(x.device, lib.sd_encode_workspace_bytes(W, T, N, H, len(plans)), "encode")
        z = torch.empty(W, N, fc.out_features, device=x.device, dtype=torch.float32)
        nv.check(lib.sd_encode(ih.handle, handles, len(plans), fc.handle, x.data_ptr(), W, T, Fdim, z.data_ptr(), final_act,
                               ws.data_ptr(), nv.PRECISIONS[precision], nv.stream_ptr(x.device)), "sd_encode")
        return z

    def forward(self, x: torch.Tensor, state=None):
        if state is not None:
            raise NotImplementedError("externally supplied recurrent state")
        return self.encode(x, nv.ACT_TANH if self.encoder_act == "tanh" else nv.ACT_NONE), None


class Decoder(nn.Module):
    def __init__(self, num_nodes: int, feature_size: int, input_size: int, hidden_size: int, output_size: int,
                 node_types: torch.Tensor = None, dec_num_layers: int = 1, dropout: float = 0., param_groups=None,
                 recurrent_arch_decoder: str = "StaticGraphGRU", **kwargs):
        super().__init__()
        if recurrent_arch_decoder != "StaticGraphGRU":
            raise NotImplementedError("only StaticGraphGRU is on the shipped path (LSTM variant is out of scope)")
        if dec_num_layers != 1:
            raise NotImplementedError("dec_num_layers != 1")
        self.param_groups, self.num_layers = param_groups, dec_num_layers
        self.if_consider_hip = kwargs["if_consider_hip"]
        self.activation_fn = nn.Tanh()
        self.recurrent_arch = recurrent_arch_decoder
        self.rnn = StaticGraphGRU(feature_size + input_size, hidden_size, num_nodes=num_nodes, num_layers=dec_num_layers,
                                  learn_influence=True, node_types=node_types, recurrent_dropout=dropout,
                                  learn_additive_graph_influence=True, clockwork=False)
        self.initial_hidden_h = StaticGraphLinear(feature_size + input_size, hidden_size, num_nodes=num_nodes, learn_influence=True, node_types=node_types)
        self.fc = StaticGraphLinear(hidden_size, output_size, num_nodes=num_nodes, learn_influence=True, node_types=node_types)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, h: torch.Tensor, z: torch.Tensor = None, ph: int = 1, state=None, precision: str = "fp32"):
        """x [Bx, >=2, N, F] (last two frames are used), h latent [B, N, L]; Bx may be B / k (samples of a
        window share their observation and are read in place).  Returns ([B, ph, N, F], x[:, -1])."""
        if state is not None:
            raise NotImplementedError("externally supplied recurrent state")
        nv.require_cuda(h, "latent")
        h = h.float().contiguous()
        x = x.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        B, N, _ = h.shape
        if B == 0 or x.shape[0] == 0:       # empty batch: empty prediction, like the reference's loop over zero rows
            return torch.empty(0, ph, N, self.fc.out_features, device=h.device, dtype=torch.float32), x[:, -1]
        if B % x.shape[0] != 0:
            raise ValueError("latent batch must be a multiple of the observation batch")
        rep = B // x.shape[0]
        cell = self.rnn.layers[0]
        gplan = cell.plan(ph)
        ih, fc = self.initial_hidden_h.plan(), self.fc.plan()
        feat = fc.out_features
        lib = nv.load()
        out = torch.empty(B, ph, N, feat, device=h.device, dtype=torch.float32)
        ws = Workspace.get(h.device, lib.sd_decode_workspace_bytes(B, N, cell.hidden_size), "decode")
        v_prev, v_last = nv.view_of(x[:, -2], rep), nv.view_of(x[:, -1], rep)
        nv.check(lib.sd_decode(ih.handle, gplan.handle, fc.handle, C.byref(v_prev), C.byref(v_last), h.data_ptr(), B, ph, feat,
                               out.data_ptr(), ws.data_ptr(), nv.PRECISIONS[precision], nv.stream_ptr(h.device)), "sd_decode")
        return out, x[:, -1]


class AutoEncoder(nn.Module):
    def __init__(self, num_nodes: int, encoder_hidden_size: int, decoder_hidden_size: int, latent_size: int,
                 node_types: torch.Tensor = None, input_size: int = 3, z_activation: str = "tanh", enc_num_layers: int = 1,
                 loss_pose_type: str = "l1", **kwargs):
        super().__init__()
        self.param_groups = [{}]
        self.latent_size, self.loss_pose_type = latent_size, loss_pose_type
        self.encoder = Encoder(num_nodes=num_nodes, input_size=input_size, hidden_size=encoder_hidden_size, output_size=latent_size,
                               node_types=node_types, enc_num_layers=enc_num_layers, recurrent_arch=kwargs["recurrent_arch_enc"])
        assert kwargs["output_size"] == input_size
        self.decoder = Decoder(num_nodes=num_nodes, input_size=latent_size, feature_size=input_size, hidden_size=decoder_hidden_size,
                               node_types=node_types, param_groups=self.param_groups, **kwargs)
        assert z_activation in ["tanh", "identity"], f"z_activation must be either 'tanh' or 'identity', but got {z_activation}"
        self._z_act = z_activation
        self.z_activation = nn.Tanh() if z_activation == "tanh" else nn.Identity()
        self.precision = "fp32"          # 'bf16x3': recurrent products on the tensor cores (fp32-grade); see get_prediction

    def forward(self, x):
        h, _ = self.encoder(x)
        return h

    def get_past_embedding(self, past, state=None, precision: str = None):
        """tanh(tanh(fc(h_T))) in one pass: the second activation is fused into the fc epilogue (autoencoder.py:51-55)."""
        enc_tanh, z_tanh = self.encoder.encoder_act == "tanh", self._z_act == "tanh"
        act = nv.ACT_TANH_TANH if (enc_tanh and z_tanh) else (nv.ACT_TANH if (enc_tanh or z_tanh) else nv.ACT_NONE)
        with torch.no_grad():
            return self.encoder.encode(past, act, precision=precision or self.precision)

    def get_embedding(self, future, state=None):
        return self.forward(future)

    def get_train_embeddings(self, y, past, state=None):
        return self.get_past_embedding(past, state=state), self.get_embedding(y, state=state)

    def decode(self, x: torch.Tensor, h: torch.Tensor, z: torch.Tensor = None, ph=1, state=None, precision: str = None):
        out, _ = self.decoder(x=x[:, -2:], h=h, z=z, ph=ph, state=state, precision=precision or self.precision)     # autoencoder.py:66-73
        return out

    def autoencode(self, y, past, ph=1, state=None):
        z_past, z = self.get_train_embeddings(y, past, state=state)
        return self.decode(past, z, z_past, ph), z_past, z

    def loss(self, y_pred, y, type=None, reduction="mean", **kwargs):
        type = self.loss_pose_type if type is None else type
        if type == "mse":
            out = (y_pred - y) ** 2
        elif type in ("l1", "L1"):
            out = (y_pred - y).abs()
        else:
            raise NotImplementedError(type)
        loss = out.sum(-1).mean(-1).mean(-1)
        return loss.mean() if reduction == "mean" else loss
