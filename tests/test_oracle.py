"""Pins the CPU oracle (PARITY UNPINNED by the reference: it ships no tests or golden vectors and
PyG is not importable here): closed-form known answers, gradcheck, agreement of the
two independent restatements, and the committed fp64 fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dense_gat, pyg_gat, synth

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _layer(Fin, C, H, concat, Fe, slope=0.2, seed=0, dtype=torch.float64):
    torch.manual_seed(seed)
    m = pyg_gat.OracleGATConv(Fin, C, heads=H, concat=concat, negative_slope=slope, edge_dim=Fe).to(dtype)
    with torch.no_grad():
        m.bias.normal_()
    return m


def test_zero_attention_vectors_give_uniform_mean():
    # att_* = 0 -> all logits 0 -> alpha = 1/N over the N-1 neighbours + the self loop -> out = mean_j P_j
    B, N, Fin, Fe, H, C = 2, 6, 5, 3, 2, 4
    bt = synth.random_complete_batch(B, N, Fin, Fe, seed=3, dtype=torch.float64)
    m = _layer(Fin, C, H, True, Fe)
    with torch.no_grad():
        m.att_src.zero_(); m.att_dst.zero_(); m.att_edge.zero_()
    out, (ei2, alpha) = m(bt.x, bt.edge_index, bt.edge_attr, return_attention_weights=True)
    assert torch.allclose(alpha, torch.full_like(alpha, 1.0 / N), atol=1e-15)
    P = (bt.x @ m.lin_src.weight.t()).view(B, N, H * C)
    expect = P.mean(dim=1, keepdim=True).expand(B, N, H * C).reshape(B * N, H * C) + m.bias
    assert torch.allclose(out, expect, atol=1e-13)


def test_two_node_closed_form():
    # N=2, H=C=1: target 0 sees edge 1->0 and its self loop whose attribute is the mean of its single
    # incoming edge, i.e. the same attribute.
    x = torch.tensor([[1.0], [3.0]], dtype=torch.float64)
    ei = torch.tensor([[0, 1], [1, 0]])
    ea = torch.tensor([[0.5], [-2.0]], dtype=torch.float64)      # edge 0: 0->1, edge 1: 1->0
    w, a_s, a_d, we, a_e, slope = 2.0, 0.3, -0.7, 1.5, 0.4, 0.2
    out, (ei2, alpha) = pyg_gat.gat_conv_edgelist(
        x, ei, ea, torch.tensor([[w]], dtype=torch.float64), torch.tensor([[[a_s]]], dtype=torch.float64),
        torch.tensor([[[a_d]]], dtype=torch.float64), torch.tensor([[we]], dtype=torch.float64),
        torch.tensor([[[a_e]]], dtype=torch.float64), None, 1, 1, True, slope, return_attention_weights=True)
    p = x[:, 0] * w
    lrelu = lambda z: z if z > 0 else slope * z
    # target 0: sources {1 (edge attr -2.0), 0 (loop attr -2.0)}
    z10 = lrelu(a_s * p[1] + a_d * p[0] + a_e * we * -2.0)
    z00 = lrelu(a_s * p[0] + a_d * p[0] + a_e * we * -2.0)
    e = np.exp([float(z10), float(z00)]); al = e / e.sum()
    assert abs(out[0, 0].item() - (al[0] * p[1] + al[1] * p[0]).item()) < 1e-14
    assert ei2.tolist() == [[0, 1, 0, 1], [1, 0, 0, 1]]           # loops appended last
    assert abs(alpha[1, 0].item() - al[0]) < 1e-15 and abs(alpha[2, 0].item() - al[1]) < 1e-15


def test_unit_slope_makes_logits_linear():
    B, N, Fin, Fe, H, C = 2, 5, 4, 3, 2, 3
    bt = synth.random_complete_batch(B, N, Fin, Fe, seed=5, dtype=torch.float64)
    m = _layer(Fin, C, H, False, Fe, slope=1.0)
    out, (ei2, alpha) = m(bt.x, bt.edge_index, bt.edge_attr, return_attention_weights=True)
    # with slope 1 the target term d_i cancels in the softmax: alpha depends on s_j + g_ij only
    with torch.no_grad():
        m.att_dst.mul_(-3.0)
    out2, (_, alpha2) = m(bt.x, bt.edge_index, bt.edge_attr, return_attention_weights=True)
    assert torch.allclose(alpha, alpha2, atol=1e-13)


def test_attention_rows_sum_to_one_and_edge_order():
    B, N = 3, 30
    bt = synth.random_complete_batch(B, N, 8, 4, seed=7, dtype=torch.float64)
    m = _layer(8, 5, 3, False, 4)
    _, (ei2, alpha) = m(bt.x, bt.edge_index, bt.edge_attr, return_attention_weights=True)
    assert ei2.shape[1] == B * N * (N - 1) + B * N
    sums = torch.zeros(B * N, 3, dtype=torch.float64).index_add_(0, ei2[1], alpha)
    assert torch.allclose(sums, torch.ones_like(sums), atol=1e-13)


def test_gradcheck_edgelist():
    B, N, Fin, Fe, H, C = 2, 4, 3, 2, 2, 3
    bt = synth.random_complete_batch(B, N, Fin, Fe, seed=11, dtype=torch.float64)
    m = _layer(Fin, C, H, False, Fe)
    params = [p.detach().clone().requires_grad_() for p in
              (m.lin_src.weight, m.att_src, m.att_dst, m.lin_edge.weight, m.att_edge, m.bias)]
    x = bt.x.clone().requires_grad_()

    def fn(x, W, a_s, a_d, We, a_e, b):
        return pyg_gat.gat_conv_edgelist(x, bt.edge_index, bt.edge_attr, W, a_s, a_d, We, a_e, b, H, C, False, 0.2)
    assert torch.autograd.gradcheck(fn, (x, *params), eps=1e-6, atol=1e-6)


@pytest.mark.parametrize("concat", [False, True])
@pytest.mark.parametrize("shape", [(4, 7, 9, 5, 3, 5), (2, 30, 12, 6, 6, 10), (3, 2, 4, 3, 1, 2)])
def test_dense_restatement_matches_edgelist(concat, shape):
    B, N, Fin, Fe, H, C = shape
    bt = synth.random_complete_batch(B, N, Fin, Fe, seed=1, dtype=torch.float64)
    m = _layer(Fin, C, H, concat, Fe)
    x = bt.x.clone().requires_grad_()
    out, (ei2, al) = m(x, bt.edge_index, bt.edge_attr, return_attention_weights=True)
    dout = torch.randn_like(out)
    out.backward(dout)
    T = dense_gat.pyg_to_dense_tile(bt.edge_attr, bt.edge_index, B, N)
    W, a_s, a_d, We, a_e, bias = [p.detach() for p in (m.lin_src.weight, m.att_src, m.att_dst,
                                                       m.lin_edge.weight, m.att_edge, m.bias)]
    fw = dense_gat.dense_forward(bt.x, T, W, a_s, a_d, We, a_e, bias, H, C, concat, 0.2)
    assert (fw["out"] - out).abs().max() < 1e-12
    assert (dense_gat.alpha_tile_to_pyg(fw["alpha"], bt.edge_index, B, N) - al).abs().max() < 1e-13
    g = dense_gat.dense_backward(fw, bt.x, T, W, a_s, a_d, We, a_e, dout, H, C, concat, 0.2, need_dx=True)
    for name, ref in (("lin_weight", m.lin_src.weight.grad), ("att_src", m.att_src.grad),
                      ("att_dst", m.att_dst.grad), ("lin_edge_weight", m.lin_edge.weight.grad),
                      ("att_edge", m.att_edge.grad), ("bias", m.bias.grad), ("x", x.grad)):
        assert (g[name] - ref).abs().max() <= 1e-12 * max(1.0, ref.abs().max().item()), name


def test_input_self_loops_are_replaced():
    # PyG removes self loops present in the input before adding its own mean-filled ones
    B, N = 2, 5
    bt = synth.random_complete_batch(B, N, 4, 3, seed=2, dtype=torch.float64)
    m = _layer(4, 3, 2, True, 3)
    ref = m(bt.x, bt.edge_index, bt.edge_attr)
    loops = torch.arange(B * N)
    ei = torch.cat([torch.stack([loops, loops]), bt.edge_index], 1)
    ea = torch.cat([torch.randn(B * N, 3, dtype=torch.float64) * 100, bt.edge_attr], 0)
    assert torch.allclose(m(bt.x, ei, ea), ref, atol=1e-13)


def test_windowed_layout_matches_appendix_b():
    N, L, T = 6, 3, 12
    vol, vv = synth.synthetic_matrices(T, N, seed=9)
    s = synth.window_sample(vol, vv, 2, L)
    assert s.x.shape == (N, N * L) and s.edge_attr.shape == (N * (N - 1), 3 * L)
    i, c, t = 4, 1, 2
    assert s.x[i, c * L + t] == np.float32(vol[2 + t, i, c])
    e = 7
    src, dst = s.edge_index[:, e].tolist()
    assert s.edge_attr[e, 0 * L + t] == np.float32(vv[2 + t, src, dst])
    assert s.edge_attr[e, 1 * L + t] == np.float32(vv[2 + t, src, src])
    assert s.edge_attr[e, 2 * L + t] == np.float32(vv[2 + t, dst, dst])
    assert s.y_x[i] == np.float32(vol[2 + L, i, i])


def test_model_layer_rules():
    # utils/models.py:86-113
    assert pyg_gat.gat_layer_plan(10, 3, [8], True) == [(10, 8, False)]
    assert pyg_gat.gat_layer_plan(10, 3, [8, 4], True) == [(10, 8, True), (24, 4, False)]
    assert pyg_gat.gat_layer_plan(10, 3, [8, 4], False) == [(10, 8, False), (8, 4, False)]
    assert pyg_gat.gat_layer_plan(10, 1, [8, 4, 2], True) == [(10, 8, True), (8, 4, True), (4, 2, False)]
    assert pyg_gat.gat_layer_plan(10, 2, [8, 4, 2], True) == [(10, 8, True), (16, 4, True), (8, 2, False)]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(path):
    """Both restatements reproduce the committed fp64 fixtures (tests/golden/make_golden.py)."""
    z = np.load(path)
    B, N, Fin, Fe, H, C, concat = [int(v) for v in z["meta"]]
    slope = float(z["slope"])
    t = lambda k: torch.from_numpy(z[k]).double()
    x, ea, ei, dout = t("x").requires_grad_(), t("edge_attr"), torch.from_numpy(z["edge_index"]), t("dout")
    params = [t(k).requires_grad_() for k in ("lin_weight", "att_src", "att_dst", "lin_edge_weight", "att_edge", "bias")]
    out, (ei2, alpha) = pyg_gat.gat_conv_edgelist(x, ei, ea, *params, H, C, bool(concat), slope,
                                                  return_attention_weights=True)
    out.backward(dout)
    assert np.abs(out.detach().numpy() - z["out"]).max() < 1e-12
    assert np.abs(alpha.detach().numpy() - z["alpha"]).max() < 1e-13
    assert (ei2.numpy() == z["edge_index_with_loops"]).all()
    for p, k in zip(params, ("lin_weight", "att_src", "att_dst", "lin_edge_weight", "att_edge", "bias")):
        ref = z["g_" + k]
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), k
    assert len(GOLDEN) >= 5


def test_multi_output_window_targets_follow_the_reference_layout():
    """CovarianceLaggedMultiOutputDataset (utils/dataset.py:380-405): y_x is the last lag's [N, K] block of the next K
    diagonals, flattened node-major; x and edge_attr are those of the single-output dataset."""
    import numpy as np
    from oracle import synth
    N, L, K = 5, 3, 4
    vol, vv = synth.synthetic_matrices(L + K + 3, N, seed=9)
    vol, vv = np.asarray(vol), np.asarray(vv)
    t0 = 2
    s1, sK = synth.window_sample(vol, vv, t0, L), synth.window_sample(vol, vv, t0, L, future_steps=K)
    assert torch.equal(s1.x, sK.x) and torch.equal(s1.edge_attr, sK.edge_attr) and torch.equal(s1.edge_index, sK.edge_index)
    y = sK.y_x.view(N, K)
    for k in range(K):
        assert torch.equal(y[:, k], torch.tensor(np.diag(vol[t0 + L + k]), dtype=torch.float))
    assert torch.equal(y[:, 0], s1.y_x)
