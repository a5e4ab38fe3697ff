"""``GATModel`` — the SpotV2Net model with the reference's constructor and ``forward(data)``
(/root/reference/utils/models.py:61-152), built on the B200 :class:`GATConv`.

Attribute and state-dict names match the reference (``gat_layers.<i>.*``, ``linear.*``,
optional ``bnorm_node`` / ``bnorm_edge``), so ``load_state_dict`` of weights saved by
5_train_SpotV2Net.py:195 works (6_results.ipynb:318).
"""
from __future__ import annotations

import sys
from typing import Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .gat_conv import GATConv, topology_from_edge_index


def gat_layer_plan(num_node_features: int, num_heads: int, dim_hidden_layers: Sequence[int], concat_heads: bool):
    """(in_channels, out_channels, concat) per layer — the rules of utils/models.py:86-113:
    a single hidden layer never concatenates; otherwise every layer but the last follows
    ``concat_heads`` and the last always averages the heads."""
    dims = list(dim_hidden_layers)
    if len(dims) == 1:
        return [(num_node_features, dims[0], False)]
    wide = bool(concat_heads) and num_heads > 1
    plan = [(num_node_features, dims[0], bool(concat_heads))]
    for i in range(1, len(dims)):
        plan.append((dims[i - 1] * num_heads if wide else dims[i - 1], dims[i],
                     False if i == len(dims) - 1 else bool(concat_heads)))
    return plan


class GATModel(nn.Module):
    def __init__(self, num_node_features, num_edge_features, num_heads, output_node_channels,
                 dim_hidden_layers=[100], dropout_att=0.0, dropout=0.0, activation='relu',
                 concat_heads=False, negative_slope=0.2, standardize=False):
        super().__init__()
        self.dropout = dropout
        self.activation = activation
        # 6_results.ipynb's analysis variant of this class calls every layer with return_attention_weights=True and keeps
        # the result in self.attention_weights; set collect_attention = True for the same behaviour
        self.collect_attention = False
        self.attention_weights = []
        self.standardize = standardize
        if self.standardize:
            self.bnorm_node = nn.BatchNorm1d(num_node_features, affine=False)
            self.bnorm_edge = nn.BatchNorm1d(num_edge_features, affine=False)
        self.gat_layers = nn.ModuleList(
            GATConv(in_channels=fi, out_channels=fo, heads=num_heads, concat=cc, dropout=dropout_att,
                    edge_dim=num_edge_features, negative_slope=negative_slope)
            for fi, fo, cc in gat_layer_plan(num_node_features, num_heads, dim_hidden_layers, concat_heads))
        self.linear = nn.Linear(list(dim_hidden_layers)[-1], output_node_channels)
        if self.activation == 'relu':
            self.a = F.relu
        elif self.activation == 'tanh':
            self.a = torch.tanh
        elif self.activation == 'sigmoid':
            self.a = torch.sigmoid
        else:
            print('Choose an available activation function')
            sys.exit()

    def set_precision(self, precision: str):
        """"fp32" (default, 1e-5 parity) or "half" (one fp16 tensor-core product in every layer's projections)."""
        for layer in self.gat_layers:
            layer.precision = precision
        return self

    def forward(self, data):
        x, edge_index, edge_attr = data.x, data.edge_index, data.edge_attr
        if self.standardize:
            if edge_attr is None:
                raise ValueError("standardize=True normalises edge_attr: collate with structured=False")
            x = self.bnorm_node(x)
            edge_attr = self.bnorm_edge(edge_attr)
        # one topology check per batch, shared by all layers (a batch produced by
        # spotv2net_b200.data carries it already)
        topo = getattr(data, "spot_topology", None)
        if topo is None:
            topo = topology_from_edge_index(edge_index, x.shape[0], getattr(data, "nodes_per_graph", None))
        if self.collect_attention:
            self.attention_weights = []
        # batches of a structured WindowDataset carry window references: the layers read the [L, N, N] windows instead of
        # edge_attr (not with standardize=True: BatchNorm changes the edge features, which then must be materialised)
        # A materialised edge_attr always wins: the caller may have edited it (train() scales it by scale_up,
        # 5_train_SpotV2Net.py:145-147), and the windows would silently ignore that.
        win = None if (self.standardize or edge_attr is not None) else getattr(data, "spot_windows", None)
        for layer in self.gat_layers:
            if self.collect_attention:
                x, att = layer(x, edge_index, edge_attr, return_attention_weights=True, topology=topo, windows=win)
                self.attention_weights.append(att)
            else:
                x = layer(x, edge_index, edge_attr, topology=topo, windows=win)
            x = self.a(x)
            if self.dropout:
                x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.linear(x)
        return x.view(-1)
