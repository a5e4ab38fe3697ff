"""Edge-list oracle: PyG 2.3.0 ``GATConv`` semantics in PyG's own op order.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``; parity unpinned by the
reference).  Every function names the PyG 2.3.0 source it restates
([PyG] = ``torch_geometric==2.3.0``, pinned at /root/reference/README.md:88)
and the reference call site that reaches it.

The op sequence is the one written out in SURVEY.md Appendix A.1; autograd
differentiates it, exactly as the reference relies on autograd
(/root/reference/5_train_SpotV2Net.py:150-159).
"""
from __future__ import annotations

import math
import sys
from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor, nn


# --------------------------------------------------------------------------- utils
def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor]):
    """[PyG] utils/loop.py::remove_self_loops — boolean-mask compaction."""
    keep = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, keep]
    if edge_attr is not None:
        edge_attr = edge_attr[keep]
    return edge_index, edge_attr


def scatter_mean_rows(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """[PyG] utils/scatter.py::scatter(reduce='mean') for dim=0.

    sum via ``scatter_add_``; count via ``scatter_add_`` of ones, clamped at 1.
    """
    count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(index.numel()))
    count = count.clamp_(min=1)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    total = src.new_zeros((dim_size,) + tuple(src.shape[1:])).scatter_add_(0, idx, src)
    return total / count.view(-1, *([1] * (src.dim() - 1)))


def add_self_loops_mean(edge_index: Tensor, edge_attr: Optional[Tensor], num_nodes: int):
    """[PyG] utils/loop.py::add_self_loops(fill_value='mean').

    Loops (i, i) for i in range(num_nodes) are appended AFTER the real edges;
    their attribute is the mean of the attributes of the edges whose target
    (row 1 of ``edge_index``) is i.
    """
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loop_index = torch.stack([loop, loop], dim=0)
    if edge_attr is not None:
        loop_attr = scatter_mean_rows(edge_attr, edge_index[1], num_nodes)
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    return torch.cat([edge_index, loop_index], dim=1), edge_attr


def segment_softmax(src: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """[PyG] utils/softmax.py::softmax (index form, dim=0).

    max over each target segment on detached values (``scatter_reduce_``
    'amax', include_self=False), exp, ``scatter_add_`` sum + 1e-16, divide.
    """
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    seg_max = src.new_full((num_nodes,) + tuple(src.shape[1:]), float("-inf"))
    seg_max = seg_max.scatter_reduce_(0, idx, src.detach(), "amax", include_self=False)
    out = (src - seg_max.index_select(0, index)).exp()
    seg_sum = src.new_zeros((num_nodes,) + tuple(src.shape[1:])).scatter_add_(0, idx, out) + 1e-16
    return out / seg_sum.index_select(0, index)


# --------------------------------------------------------------------------- functional
def gat_conv_edgelist(
    x: Tensor,
    edge_index: Tensor,
    edge_attr: Optional[Tensor],
    lin_weight: Tensor,
    att_src: Tensor,
    att_dst: Tensor,
    lin_edge_weight: Optional[Tensor],
    att_edge: Optional[Tensor],
    bias: Optional[Tensor],
    heads: int,
    out_channels: int,
    concat: bool,
    negative_slope: float = 0.2,
    dropout: float = 0.0,
    training: bool = False,
    add_self_loops: bool = True,
    return_attention_weights: bool = False,
    dropout_mask: Optional[Tensor] = None,
):
    """[PyG] nn/conv/gat_conv.py::GATConv.forward + edge_update + message.

    Reached from /root/reference/utils/models.py:146 (``l(x, edge_index,
    edge_attr)``).  ``flow='source_to_target'``: j = edge_index[0] is the
    source, i = edge_index[1] the target and softmax group; aggregation 'add'.
    """
    assert x.dim() == 2, "Static graphs not supported in 'GATConv'"
    H, C = heads, out_channels
    n = x.shape[0]
    xs = F.linear(x, lin_weight).view(n, H, C)                # lin_src (== lin_dst)
    alpha_src = (xs * att_src.view(1, H, C)).sum(dim=-1)       # [n, H]
    alpha_dst = (xs * att_dst.view(1, H, C)).sum(dim=-1)

    if add_self_loops:
        edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
        edge_index, edge_attr = add_self_loops_mean(edge_index, edge_attr, n)

    src, dst = edge_index[0], edge_index[1]
    a = alpha_src.index_select(0, src) + alpha_dst.index_select(0, dst)
    if edge_attr is not None and lin_edge_weight is not None:
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.view(-1, 1)
        e = F.linear(edge_attr, lin_edge_weight).view(-1, H, C)
        a = a + (e * att_edge.view(1, H, C)).sum(dim=-1)
    a = F.leaky_relu(a, negative_slope)
    a = segment_softmax(a, dst, n)
    if dropout_mask is not None:
        # test hook: F.dropout with a GIVEN Bernoulli draw ([E', H] of 0/1, PyG edge order with the self loops
        # last) instead of torch's generator, so another implementation's mask can be replayed exactly
        alpha = a * dropout_mask.to(a.dtype) / (1.0 - dropout)
    else:
        alpha = F.dropout(a, p=dropout, training=training)

    msg = alpha.unsqueeze(-1) * xs.index_select(0, src)        # [E', H, C]
    out = xs.new_zeros(n, H, C).index_add_(0, dst, msg)
    out = out.reshape(n, H * C) if concat else out.mean(dim=1)
    if bias is not None:
        out = out + bias
    if return_attention_weights:
        return out, (edge_index, alpha)
    return out


# --------------------------------------------------------------------------- modules
def glorot_(t: Tensor) -> Tensor:
    """[PyG] nn/inits.py::glorot — U(±sqrt(6 / (size[-2] + size[-1])))."""
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-bound, bound)


class _GlorotLinear(nn.Module):
    """[PyG] nn/dense/linear.py::Linear(bias=False, weight_initializer='glorot')."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight)

    def forward(self, x):
        return F.linear(x, self.weight)


class OracleGATConv(nn.Module):
    """Module form with PyG 2.3.0 parameter names and init draw order
    (SURVEY.md Appendix A.5): ``lin_src`` and ``lin_edge`` draw at
    construction, then ``reset_parameters`` draws lin_src, lin_dst (the same
    module again), lin_edge, att_src, att_dst, att_edge; bias = 0."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, edge_dim=None, fill_value="mean", bias=True):
        super().__init__()
        assert fill_value == "mean", "the reference never overrides fill_value"
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.edge_dim = add_self_loops, edge_dim
        self.lin_src = _GlorotLinear(in_channels, heads * out_channels)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if edge_dim is not None:
            self.lin_edge = _GlorotLinear(edge_dim, heads * out_channels)
            self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_src.reset_parameters()
        self.lin_dst.reset_parameters()
        if self.lin_edge is not None:
            self.lin_edge.reset_parameters()
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if self.att_edge is not None:
            glorot_(self.att_edge)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x, edge_index, edge_attr=None, size=None, return_attention_weights=None, dropout_mask=None):
        return gat_conv_edgelist(
            x, edge_index, edge_attr, self.lin_src.weight, self.att_src, self.att_dst,
            None if self.lin_edge is None else self.lin_edge.weight, self.att_edge, self.bias,
            self.heads, self.out_channels, self.concat, self.negative_slope, self.dropout,
            self.training, self.add_self_loops, bool(return_attention_weights), dropout_mask)


def gat_layer_plan(num_node_features: int, num_heads: int, dim_hidden_layers: Sequence[int],
                   concat_heads: bool):
    """Layer-construction rules of /root/reference/utils/models.py:86-113 as
    a list of (in_channels, out_channels, concat) triples."""
    dims = list(dim_hidden_layers)
    widen = bool(concat_heads) and num_heads > 1
    if len(dims) == 1:
        return [(num_node_features, dims[0], False)]
    plan = [(num_node_features, dims[0], bool(concat_heads))]
    for i in range(len(dims) - 1):
        last = (i + 1 == len(dims) - 1)
        fan_in = dims[i] * num_heads if widen else dims[i]
        plan.append((fan_in, dims[i + 1], False if last else bool(concat_heads)))
    return plan


class OracleGATModel(nn.Module):
    """Twin of the reference ``GATModel`` (/root/reference/utils/models.py:61-152)
    built on :class:`OracleGATConv`.  Same ctor arguments, same attribute and
    state-dict names, same ``forward(data)`` reading ``data.x/edge_index/edge_attr``."""

    def __init__(self, num_node_features, num_edge_features, num_heads, output_node_channels,
                 dim_hidden_layers=(100,), dropout_att=0.0, dropout=0.0, activation="relu",
                 concat_heads=False, negative_slope=0.2, standardize=False):
        super().__init__()
        self.dropout, self.activation, self.standardize = dropout, activation, standardize
        if standardize:
            self.bnorm_node = nn.BatchNorm1d(num_node_features, affine=False)
            self.bnorm_edge = nn.BatchNorm1d(num_edge_features, affine=False)
        self.gat_layers = nn.ModuleList([
            OracleGATConv(fi, fo, heads=num_heads, concat=cc, dropout=dropout_att,
                          edge_dim=num_edge_features, negative_slope=negative_slope)
            for fi, fo, cc in gat_layer_plan(num_node_features, num_heads, dim_hidden_layers, concat_heads)])
        self.linear = nn.Linear(list(dim_hidden_layers)[-1], output_node_channels)
        acts = {"relu": F.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}
        if activation not in acts:
            print("Choose an available activation function")
            sys.exit()
        self.a = acts[activation]

    def forward(self, data):
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        if self.standardize:
            x = self.bnorm_node(x)
            edge_attr = self.bnorm_edge(edge_attr)
        for layer in self.gat_layers:
            x = self.a(layer(x, edge_index, edge_attr))
            if self.dropout:
                x = F.dropout(x, p=self.dropout, training=self.training)
        return self.linear(x).view(-1)
