"""Dense-tile / folded oracle: what the CUDA kernels implement, stage by stage.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Independent restatement of
the same layer as :mod:`oracle.pyg_gat` for batches of identical complete
graphs (the only topology the reference's default dataset produces,
/root/reference/utils/dataset.py:216-226), using the two algebraic facts of
SURVEY.md §0.6 / Appendix A.2:

* edge-term fold:  a_edge,h . (W_e,h e_ij)  ==  e_ij . v_h,   v_h = W_e,h^T a_edge,h
* the self-loop PyG synthesises with fill_value='mean' contributes the
  row-mean of the other N-1 edge terms of that target.

and the "augmented projection" the kernels use:

* s = P a_src == x (W_h^T a_src,h) =: x u_src,h ; likewise d.  The 2H vectors
  u are appended to W as extra output rows, so the projection GEMM emits
  P_aug = [P | s | d] and the weight-gradient GEMM consumes dP_aug = [dP | ds | dd].

The backward here is the hand-derived closed form,
NOT autograd, so that agreement with autograd of the edge-list oracle pins
both.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor


def complete_edge_table(edge_index_one_graph: Tensor, N: int) -> Tensor:
    """rows -> (target i, source j) for one graph's local edge_index [2, R]."""
    return torch.stack([edge_index_one_graph[1], edge_index_one_graph[0]], dim=1)


def pyg_to_dense_tile(edge_attr: Tensor, edge_index: Tensor, B: int, N: int) -> Tensor:
    """[B*R, Fe] in PyG concatenated order -> T[B, N, N, Fe] with
    T[b, i, j] = attribute of edge j -> i (diagonal left at zero)."""
    R = edge_index.shape[1] // B
    Fe = edge_attr.shape[1]
    local = edge_index[:, :R]
    T = edge_attr.new_zeros(B, N, N, Fe)
    T[:, local[1], local[0], :] = edge_attr.view(B, R, Fe)
    return T


def fold_params(W: Tensor, a_src: Tensor, a_dst: Tensor, W_e: Optional[Tensor],
                a_edge: Optional[Tensor], H: int, C: int):
    """W_aug [H*C + 2H, F] = [W ; u_src ; u_dst],  v [H, Fe]."""
    F_in = W.shape[1]
    Wh = W.view(H, C, F_in)
    u_src = torch.einsum("hcf,hc->hf", Wh, a_src.view(H, C))
    u_dst = torch.einsum("hcf,hc->hf", Wh, a_dst.view(H, C))
    W_aug = torch.cat([W, u_src, u_dst], dim=0)
    v = None
    if W_e is not None:
        v = torch.einsum("hcf,hc->hf", W_e.view(H, C, -1), a_edge.view(H, C))
    return W_aug, v


def dense_forward(x: Tensor, T: Optional[Tensor], W: Tensor, a_src: Tensor, a_dst: Tensor,
                  W_e: Optional[Tensor], a_edge: Optional[Tensor], bias: Optional[Tensor],
                  H: int, C: int, concat: bool, slope: float) -> Dict[str, Tensor]:
    """Appendix A.2.  x [B*N, F]; T [B, N, N, Fe] target-major (diag ignored)."""
    B, N = T.shape[0], T.shape[1]          # pass a zero tile with W_e=None when there are no edge features
    W_aug, v = fold_params(W, a_src, a_dst, W_e, a_edge, H, C)
    P_aug = x @ W_aug.t()
    HC = H * C
    n = x.shape[0]
    P =P_aug[:, :HC].reshape(B, N, H, C)
    s = P_aug[:, HC:HC + H].reshape(B, N, H)
    d = P_aug[:, HC + H:HC + 2 * H].reshape(B, N, H)
    if v is not None:
        g = torch.einsum("bijf,hf->bijh", T, v)
        eye = torch.eye(N, dtype=torch.bool, device=x.device).view(1, N, N, 1)
        g = g.masked_fill(eye, 0.0)
        g_loop = g.sum(dim=2) / max(N - 1, 1)                       # row mean of the real edges
        g = g + torch.diag_embed(g_loop.permute(0, 2, 1)).permute(0, 2, 3, 1)
    else:
        g = x.new_zeros(B, N, N, H)
    z = s.unsqueeze(1) + d.unsqueeze(2) + g                         # z[b,i,j,h] = s_j + d_i + g_ij
    l = torch.where(z > 0, z, z * slope)
    alpha = torch.softmax(l, dim=2)
    O = torch.einsum("bijh,bjhc->bihc", alpha, P)
    out = O.reshape(n, HC) if concat else O.mean(dim=2).reshape(n, C)
    if bias is not None:
        out = out + bias
    return dict(W_aug=W_aug, v=v, P_aug=P_aug, P=P, s=s, d=d, g=g, z=z, alpha=alpha, O=O, out=out)


def dense_backward(fw: Dict[str, Tensor], x: Tensor, T: Optional[Tensor], W: Tensor, a_src: Tensor,
                   a_dst: Tensor, W_e: Optional[Tensor], a_edge: Optional[Tensor], dout: Tensor,
                   H: int, C: int, concat: bool, slope: float, need_dx: bool = False):
    """Appendix A.3 in the augmented form: attention backward produces
    dP_aug = [dP | ds | dd]; the GEMM produces dW_aug = dP_aug^T x; the unfold
    step distributes the 2H extra rows back onto W, a_src, a_dst."""
    alpha, P, z = fw["alpha"], fw["P"], fw["z"]
    B, N = alpha.shape[0], alpha.shape[1]
    HC = H * C
    n = B * N
    dO = dout.view(B, N, H, C) if concat else (dout.view(B, N, 1, C) / H).expand(B, N, H, C)
    dalpha = torch.einsum("bihc,bjhc->bijh", dO, P)
    row = (alpha * dalpha).sum(dim=2, keepdim=True)
    dl = alpha * (dalpha - row)
    dz = dl * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, slope))
    ds = dz.sum(dim=1)                                              # over targets i -> [B, j, H]
    dd = dz.sum(dim=2)                                              # over sources j -> [B, i, H]
    dP = torch.einsum("bijh,bihc->bjhc", alpha, dO)
    dP_aug = torch.cat([dP.reshape(n, HC), ds.reshape(n, H), dd.reshape(n, H)], dim=1)
    dW_aug = dP_aug.t() @ x
    grads = dict(dP_aug=dP_aug, dW_aug=dW_aug)
    # gradient through the self-loop mean fill
    dv = None
    if fw["v"] is not None:
        diag = torch.diagonal(dz, dim1=1, dim2=2).permute(0, 2, 1)  # [B, N, H] = dz_ii
        dzp = dz + diag.unsqueeze(2) / max(N - 1, 1)
        eye = torch.eye(N, dtype=torch.bool, device=x.device).view(1, N, N, 1)
        dzp = dzp.masked_fill(eye, 0.0)
        dv = torch.einsum("bijh,bijf->hf", dzp, T)
        grads["dzp"] = dzp
    grads["dv"] = dv
    grads.update(unfold_grads(dW_aug, dv, W, a_src, a_dst, W_e, a_edge, H, C))
    grads["bias"] = dout.sum(dim=0)
    if need_dx:
        grads["x"] = dP_aug @ fw["W_aug"]
    return grads


def unfold_grads(dW_aug: Tensor, dv: Optional[Tensor], W: Tensor, a_src: Tensor, a_dst: Tensor,
                 W_e: Optional[Tensor], a_edge: Optional[Tensor], H: int, C: int):
    HC = H * C
    F_in = W.shape[1]
    dU_src, dU_dst = dW_aug[HC:HC + H], dW_aug[HC + H:HC + 2 * H]   # [H, F]
    Wh = W.view(H, C, F_in)
    dW = dW_aug[:HC].view(H, C, F_in) \
        + a_src.view(H, C, 1) * dU_src.view(H, 1, F_in) \
        + a_dst.view(H, C, 1) * dU_dst.view(H, 1, F_in)
    out = dict(
        lin_weight=dW.reshape(HC, F_in),
        att_src=torch.einsum("hcf,hf->hc", Wh, dU_src).view(1, H, C),
        att_dst=torch.einsum("hcf,hf->hc", Wh, dU_dst).view(1, H, C),
    )
    if dv is not None:
        out["lin_edge_weight"] = (a_edge.view(H, C, 1) * dv.view(H, 1, -1)).reshape(HC, -1)
        out["att_edge"] = torch.einsum("hcf,hf->hc", W_e.view(H, C, -1), dv).view(1, H, C)
    return out


def alpha_tile_to_pyg(alpha: Tensor, edge_index: Tensor, B: int, N: int) -> Tensor:
    """Appendix A.4: alpha[B,N,N,H] -> PyG order [B*R + B*N, H] (real edges in
    batch order, then the n self-loops in node order)."""
    R = edge_index.shape[1] // B
    local = edge_index[:, :R]
    real = alpha[:, local[1], local[0], :].reshape(B * R, -1)
    loops = torch.diagonal(alpha, dim1=1, dim2=2).permute(0, 2, 1).reshape(B * N, -1)
    return torch.cat([real, loops], dim=0)
