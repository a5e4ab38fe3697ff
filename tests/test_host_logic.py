"""Host-side mirror of the reference interface: constructor rules, state-dict names, init."""
import numpy as np
import pytest
import torch

import spotv2net_b200 as sv
from oracle import pyg_gat, synth


@pytest.mark.parametrize("hidden", [[500], [64, 32], [50, 25, 12]])
@pytest.mark.parametrize("concat_heads", [True, False])
@pytest.mark.parametrize("heads", [1, 3])
def test_ctor_matrix_and_state_dict_parity(hidden, concat_heads, heads):
    kw = dict(num_node_features=60, num_edge_features=9, num_heads=heads, output_node_channels=1,
              dim_hidden_layers=hidden, concat_heads=concat_heads, negative_slope=0.1)
    ours, ref = sv.GATModel(**kw), pyg_gat.OracleGATModel(**kw)
    so, sr = ours.state_dict(), ref.state_dict()
    assert list(so.keys()) == list(sr.keys())
    assert all(so[k].shape == sr[k].shape for k in so)
    assert sv.gat_layer_plan(60, heads, hidden, concat_heads) == pyg_gat.gat_layer_plan(60, heads, hidden, concat_heads)
    ours.load_state_dict(sr)                      # weights saved by the reference load by name
    for k in so:
        assert torch.equal(ours.state_dict()[k], sr[k])
    for layer, (fi, fo, cc) in zip(ours.gat_layers, sv.gat_layer_plan(60, heads, hidden, concat_heads)):
        assert (layer.in_channels, layer.out_channels, layer.concat, layer.heads) == (fi, fo, cc, heads)
        assert layer.negative_slope == 0.1 and layer.edge_dim == 9 and layer.lin_dst is layer.lin_src


def test_default_config_parameter_count():
    m = sv.GATModel(1260, 126, 6, 1, [500], concat_heads=True)          # config/GNN_param.yaml:26-38
    assert sum(p.numel() for p in m.parameters()) == 4_168_001
    assert m.gat_layers[0].concat is False                              # single layer never concatenates


def test_init_follows_pyg_draw_order():
    torch.manual_seed(123)
    a = sv.GATConv(12, 5, heads=3, edge_dim=4)
    torch.manual_seed(123)
    b = pyg_gat.OracleGATConv(12, 5, heads=3, edge_dim=4)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    bound = (6.0 / (12 + 15)) ** 0.5
    assert a.lin_src.weight.abs().max() <= bound and torch.all(a.bias == 0)


def test_pyg25_checkpoint_alias():
    a = sv.GATConv(6, 4, heads=2, edge_dim=3)
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    w = sd.pop("lin_src.weight"); sd.pop("lin_dst.weight")
    sd["lin.weight"] = w + 1
    a.load_state_dict(sd)
    assert torch.equal(a.lin_src.weight, w + 1)


def test_unknown_activation_exits_like_the_reference():
    with pytest.raises(SystemExit):
        sv.GATModel(8, 3, 2, 1, [4], activation="gelu")


def test_unsupported_layer_options_fail_loudly():
    with pytest.raises(sv.SpotV2Error):
        sv.GATConv(8, 4, add_self_loops=False)
    with pytest.raises(sv.SpotV2Error):
        sv.GATConv((8, 8), 4)


def test_edge_index_order_matches_reference_dataset():
    for N in (2, 5, 30):
        assert torch.equal(sv.complete_graph_edge_index(N), synth.complete_graph_edge_index(N))


def test_projection_format_choice_and_padded_sizes():
    """Host-side rule for spotv2_gat_desc.p_format (gat_conv.pair_format_applies / _desc) and the sizes that follow from it
    (no GPU: spotv2_gat_n_aug / spotv2_gat_head_pitch are plain host functions of the descriptor)."""
    import ctypes as C
    from spotv2net_b200 import _lib, gat_conv
    ok = gat_conv.pair_format_applies
    assert ok(30, 126, 500, 0, 0) and ok(30, 126, 256, 3, 2, concat=True) and ok(1, 0, 4, 0, 0)
    assert not ok(33, 126, 500, 0, 0)              # several CTAs per graph: fp32 P_aug
    assert not ok(30, 126, 500, 1, 0)              # CUDA-core GEMM
    assert not ok(30, 126, 500, 0, 1)              # phase-serial backward
    assert not ok(30, 126, 502, 0, 0)              # dout tiles by TMA need C % 4 == 0
    assert not ok(30, 126, 36, 0, 0, concat=True)  # concat layers: head blocks of dout start on 16-byte boundaries only if C % 8 == 0
    assert not ok(30, 400, 500, 0, 0) and not ok(30, 126, 1028, 0, 0)
    lib = _lib.load()
    topo = gat_conv.Topology(4, 30, 870, None, False)
    d1 = gat_conv._desc(topo, 1260, 126, 6, 500, False, 0.2)
    assert d1.p_format == 1 and lib.spotv2_gat_head_pitch(C.byref(d1)) == 504 and lib.spotv2_gat_n_aug(C.byref(d1)) == 6 * 504 + 12
    d0 = gat_conv._desc(topo, 1260, 126, 6, 500, False, 0.2, p_format=0)
    assert d0.p_format == 0 and lib.spotv2_gat_head_pitch(C.byref(d0)) == 500 and lib.spotv2_gat_n_aug(C.byref(d0)) == 3012
    dc = gat_conv._desc(topo, 1260, 126, 8, 256, True, 0.2, gemm_algo=3)
    assert dc.p_format == 1 and lib.spotv2_gat_head_pitch(C.byref(dc)) == 256
    # the library's own answer adds the kernels' shared-memory plans to the shape rules
    fits = lambda N, Fe, H, Cc, cat: lib.spotv2_gat_pair_format_supported(C.byref(_lib.GatDesc(8, N, 64, Fe, H, Cc, N * (N - 1), cat, 0.2,
                                                                                               lib.spotv2_gat_ldp(H, Cc), 0, 0)))
    assert fits(30, 126, 6, 500, 0) == 1 and fits(30, 126, 8, 256, 1) == 1 and fits(32, 384, 8, 512, 0) == 1
    assert fits(32, 512, 8, 256, 0) == 0 and fits(2, 3, 1, 2, 1) == 0 and fits(33, 126, 6, 500, 0) == 0
    bad = _lib.GatDesc(4, 40, 1260, 126, 6, 500, 1560, 0, 0.2, 3012, 0, 0, 0.0, 0, 0, 0, 1)
    assert lib.spotv2_gat_workspace_bytes(C.byref(bad), None, None, None) != 0 and b"p_format 1" in lib.spotv2_last_error()
