"""Equivariant layers of the 2-FGNN, backed by libfgnn_b200 CUDA kernels.

Same constructors, parameter names and state-dict keys as the reference models/layers.py
(hot subset: GraphNorm :47-69, normalize :71-80, MlpBlock_Real :109-131, Concat :145-146,
Identity :151-152, Matmul :161-162, ColumnMaxPooling :194-203).  forward() accepts the
reference's input types -- a (B,C,N,N) tensor or a MaskedTensor -- but runs exclusively on
CUDA through the C ABI; CPU tensors raise (no fallback).
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from .. import _ops
from ..maskedtensors.maskedtensor import MaskedTensor, dispatch_cat


def _unwrap(x):
    """-> (plain tensor, int32 device sizes or None, rewrap(tensor, names))."""
    if isinstance(x, MaskedTensor):
        names = x.tensor.names
        masks = x.mask_dict
        plain = x.tensor.rename(None)
        n_dev = x.sizes_i32(plain.device)

        def rewrap(t, out_names=names):
            keep = {k: v for k, v in masks.items() if k in out_names}
            return MaskedTensor(t.rename(*out_names), keep, adjust_mask=False, apply_mask=False)._inherit_sizes(x)
        return plain, n_dev, rewrap
    return x, None, (lambda t, out_names=None: t)


def normalize(b, constant_n_vertices=True, eps=1e-05):
    """(b - mean) / (2*sqrt(n*(var+eps))) over each (graph, channel) plane; n is the padded size for
    dense batches and the per-graph size for MaskedTensors (reference layers.py:71-80)."""
    plain, n_dev, rewrap = _unwrap(b)
    if not constant_n_vertices and n_dev is None:
        raise TypeError("constant_n_vertices=False needs a MaskedTensor input")
    return rewrap(_ops.graphnorm_fwd(plain, n_dev, None, None, eps, constant_n_vertices))


class GraphNorm(nn.Module):
    def __init__(self, features, constant_n_vertices=True, elementwise_affine=True, eps=1e-05,
                 device=None, dtype=None):
        super().__init__()
        self.constant_n_vertices = constant_n_vertices
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        self.features = (1, features, 1, 1)
        if elementwise_affine:
            self.weight = Parameter(torch.ones(self.features, device=device, dtype=dtype))
            self.bias = Parameter(torch.zeros(self.features, device=device, dtype=dtype))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def reset_parameters(self) -> None:
        if self.elementwise_affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def forward(self, b):
        plain, n_dev, rewrap = _unwrap(b)
        return rewrap(_ops.graphnorm_fwd(plain, n_dev, self.weight, self.bias, self.eps, self.constant_n_vertices))


def _init_weights(layer):
    """xavier_uniform weights, zero bias (reference layers.py:134-142)."""
    nn.init.xavier_uniform_(layer.weight)
    if layer.bias is not None:
        nn.init.zeros_(layer.bias)


class MlpBlock_Real(nn.Module):
    """depth_of_mlp 1x1 convolutions with ReLU between them, then GraphNorm -- one fused CUDA call.

    Parameters live in `convs` (nn.Conv2d, weight (Co,Ci,1,1)) and `gn` exactly as in the
    reference, so checkpoints load unchanged."""

    def __init__(self, in_features, out_features, depth_of_mlp, activation_fn=F.relu,
                 constant_n_vertices=True):
        super().__init__()
        if activation_fn is not F.relu and activation_fn is not torch.relu:
            raise NotImplementedError("fgnn_b200 fuses ReLU into the conv chain; other activations are unsupported")
        self.activation = activation_fn
        self.depth_mlp = depth_of_mlp
        self.cst_vertices = constant_n_vertices
        self.convs = nn.ModuleList()
        width = in_features
        for _ in range(depth_of_mlp):
            conv = nn.Conv2d(width, out_features, kernel_size=1, padding=0, bias=True)
            _init_weights(conv)
            self.convs.append(conv)
            width = out_features
        self.gn = GraphNorm(out_features, constant_n_vertices=constant_n_vertices)

    def forward(self, inputs):
        plain, n_dev, rewrap = _unwrap(inputs)
        ws = [c.weight for c in self.convs]
        bs = [c.bias for c in self.convs]
        y = _ops.MlpFunction.apply(plain, n_dev, self.gn.eps, len(ws), bool(self.cst_vertices), self.gn.weight,
                                   self.gn.bias, *ws, *bs)
        return rewrap(y)


class Concat(nn.Module):
    def forward(self, *xs):
        if any(isinstance(x, MaskedTensor) for x in xs):
            return dispatch_cat(xs, dim=1)
        return torch.cat(xs, dim=1)


class Diag(nn.Module):
    def forward(self, xs):
        return torch.diag_embed(xs)


class Identity(namedtuple('Identity', [])):
    def __call__(self, x):
        return x


class Permute(namedtuple('Permute', [])):
    def __call__(self, x):
        return x.permute(0, 2, 1)


class Add(nn.Module):
    def forward(self, xs1, xs2):
        return torch.add(xs1, xs2)


class Matmul(nn.Module):
    """Per-(graph, channel) N x N matrix product."""

    def forward(self, xs1, xs2):
        a, n_dev, rewrap = _unwrap(xs1)
        b, n_dev2, _ = _unwrap(xs2)
        if (n_dev is None) != (n_dev2 is None):
            raise TypeError("Matmul: both operands must be masked or both dense")
        return rewrap(_ops.MatmulFunction.apply(a, b, n_dev))


class ColumnMaxPooling(nn.Module):
    """(bs, features, n, n) -> (bs, features, n): max over the last dim."""

    def forward(self, x):
        plain, n_dev, rewrap = _unwrap(x)
        out = _ops.ColMaxFunction.apply(plain, n_dev)
        if isinstance(x, MaskedTensor):
            return rewrap(out, x.tensor.names[:-1])
        return out
