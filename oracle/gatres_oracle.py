"""CPU oracle for the GATRes message-passing hot path.  TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  The arithmetic of this path lives in PyTorch Geometric
(`torch_geometric`, ">=2.3" per /root/reference/README.md:32), which is neither
vendored in the reference tree nor installable in this image, and the reference
ships no tests, golden vectors or checkpoints.  This file is therefore a
*restatement* of the published PyG operator semantics,
anchored on the reference's own call sites:

  * model structure       gnn_pressure_estimation/GraphModels.py:454-494
  * GATConv construction  GraphModels.py:458-459  (heads=2 concat / heads=1 mean)
  * SimpleConv(mean)      GraphModels.py:460,466
  * Linear(1,nc)/(nc,1)   GraphModels.py:477,484
  * caller (mask, loss)   gnn_pressure_estimation/train.py:171-185,
                          gnn_pressure_estimation/utils/auxil.py:143-182
  * collation             train.py:302 (PyG DataLoader), SURVEY.md §A.5

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product path
(`gnn_pressure_estimation_b200`) never does and has no CPU fallback.

Everything is plain PyTorch on CPU tensors, dtype-generic (fp32 for parity,
fp64 to arbitrate tolerance questions).  Three independent formulations of the
GAT aggregation are provided so the oracle can be checked against itself:
edge-list scatter (the op sequence PyG issues), dense adjacency, and a pure
Python per-row loop.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

NEG_SLOPE = 0.2          # GATConv default negative_slope [ext PyG]
SOFTMAX_EPS = 1e-16      # torch_geometric.utils.softmax adds this to the denominator


# --------------------------------------------------------------------------- #
# edge-list helpers
# --------------------------------------------------------------------------- #
def rewrite_edges(edge_index: Tensor, num_nodes: int) -> Tensor:
    """remove_self_loops + add_self_loops as GATConv.forward does on every call:
    existing (n,n) edges are dropped (order of the rest preserved) and one
    self-loop per node is appended at the END of the list."""
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    loops = torch.arange(num_nodes, dtype=ei.dtype)
    return torch.cat([ei, torch.stack([loops, loops])], dim=1)


def segment_softmax(a: Tensor, index: Tensor, num_segments: int) -> Tensor:
    """torch_geometric.utils.softmax: max on detached values, exp, sum + 1e-16."""
    mx = torch.full((num_segments,) + a.shape[1:], float("-inf"), dtype=a.dtype)
    mx = mx.scatter_reduce(0, index.view(-1, *[1] * (a.dim() - 1)).expand_as(a), a.detach(),
                           reduce="amax", include_self=True)
    p = (a - mx.index_select(0, index)).exp()
    den = torch.zeros((num_segments,) + a.shape[1:], dtype=a.dtype).index_add_(0, index, p) + SOFTMAX_EPS
    return p / den.index_select(0, index)


# --------------------------------------------------------------------------- #
# GATConv / SimpleConv restatements
# --------------------------------------------------------------------------- #
def gat_conv(x: Tensor, edge_index: Tensor, weight: Tensor, att_src: Tensor, att_dst: Tensor,
             bias: Optional[Tensor], heads: int, concat: bool,
             return_alpha: bool = False):
    """GATConv(in, C, heads, concat).forward(x, edge_index)  (SURVEY §A.2).

    weight [H*C, in]; att_* [1,H,C]; bias [H*C] (concat) or [C] (mean).
    Follows PyG's op order: projection GEMM, per-node scores, edge rewrite,
    per-edge gather + LeakyReLU, segment softmax, [E',H,C] messages, index_add_.
    """
    M = x.size(0)
    H = heads
    C = weight.size(0) // H
    h = F.linear(x, weight).view(M, H, C)
    s_src = (h * att_src).sum(-1)
    s_dst = (h * att_dst).sum(-1)
    ei = rewrite_edges(edge_index, M)
    j, i = ei[0], ei[1]
    a = F.leaky_relu(s_src.index_select(0, j) + s_dst.index_select(0, i), NEG_SLOPE)
    alpha = segment_softmax(a, i, M)
    msg = alpha.unsqueeze(-1) * h.index_select(0, j)
    out = torch.zeros(M, H, C, dtype=x.dtype).index_add_(0, i, msg)
    out = out.reshape(M, H * C) if concat else out.mean(dim=1)
    if bias is not None:
        out = out + bias
    if return_alpha:
        return out, (ei, alpha)
    return out


def simple_conv_mean(x: Tensor, edge_index: Tensor) -> Tensor:
    """SimpleConv(aggr='mean').forward(x, edge_index)  (SURVEY §A.3): mean of
    in-neighbour rows over the ORIGINAL edge list (no self-loops added);
    count clamped to >= 1 so isolated nodes give 0."""
    M = x.size(0)
    j, i = edge_index[0], edge_index[1]
    tot = torch.zeros_like(x).index_add_(0, i, x.index_select(0, j))
    cnt = torch.zeros(M, dtype=x.dtype).index_add_(0, i, torch.ones(i.numel(), dtype=x.dtype))
    return tot / cnt.clamp(min=1).unsqueeze(-1)


def gat_conv_dense(x, edge_index, weight, att_src, att_dst, bias, heads, concat):
    """Same operator through a dense [M,M] adjacency (independent formulation)."""
    M = x.size(0)
    H = heads
    C = weight.size(0) // H
    h = (x @ weight.t()).view(M, H, C)
    s_src = torch.einsum("mhc,hc->mh", h, att_src[0])
    s_dst = torch.einsum("mhc,hc->mh", h, att_dst[0])
    ei = rewrite_edges(edge_index, M)
    cnt = torch.zeros(M, M, dtype=x.dtype)                       # multiplicity of edge j->i at [i, j]
    cnt.index_put_((ei[1], ei[0]), torch.ones(ei.size(1), dtype=x.dtype), accumulate=True)
    z = F.leaky_relu(s_dst.unsqueeze(1) + s_src.unsqueeze(0), NEG_SLOPE)   # [i, j, H]
    z = z.masked_fill((cnt == 0).unsqueeze(-1), float("-inf"))
    p = (z - z.max(dim=1, keepdim=True).values).exp() * cnt.unsqueeze(-1)
    alpha = p / (p.sum(dim=1, keepdim=True) + SOFTMAX_EPS)
    out = torch.einsum("ijh,jhc->ihc", alpha, h)
    out = out.reshape(M, H * C) if concat else out.mean(dim=1)
    return out + bias if bias is not None else out


def gat_conv_rowloop(x, edge_index, weight, att_src, att_dst, bias, heads, concat):
    """Pure-Python per-target-row loop (small cases only)."""
    M = x.size(0)
    H = heads
    C = weight.size(0) // H
    h = (x @ weight.t()).view(M, H, C)
    ei = rewrite_edges(edge_index, M)
    in_edges = [[] for _ in range(M)]
    for e in range(ei.size(1)):
        in_edges[int(ei[1, e])].append(int(ei[0, e]))
    out = torch.zeros(M, H, C, dtype=x.dtype)
    for i in range(M):
        for hh in range(H):
            sd = float((h[i, hh] * att_dst[0, hh]).sum())
            zs = []
            for jn in in_edges[i]:
                zz = float((h[jn, hh] * att_src[0, hh]).sum()) + sd
                zs.append(zz if zz > 0 else NEG_SLOPE * zz)
            mx = max(zs)
            ps = [math.exp(v - mx) for v in zs]
            den = sum(ps) + SOFTMAX_EPS
            for jn, pv in zip(in_edges[i], ps):
                out[i, hh] += (pv / den) * h[jn, hh]
    out = out.reshape(M, H * C) if concat else out.mean(dim=1)
    return out + bias if bias is not None else out


def gat_conv_backward_manual(x, edge_index, weight, att_src, att_dst, heads, concat, grad_out):
    """Hand-derived backward of gat_conv, the formulas the CUDA
    backward implements.  Returns dict(dx, dW, datt_src, datt_dst, dbias)."""
    M = x.size(0)
    H = heads
    C = weight.size(0) // H
    h = (x @ weight.t()).view(M, H, C)
    s_src = (h * att_src).sum(-1)
    s_dst = (h * att_dst).sum(-1)
    ei = rewrite_edges(edge_index, M)
    j, i = ei[0], ei[1]
    z = s_src[j] + s_dst[i]
    a = F.leaky_relu(z, NEG_SLOPE)
    alpha = segment_softmax(a, i, M)
    g = grad_out.view(M, H, C) if concat else (grad_out / H).unsqueeze(1).expand(M, H, C)
    dbias = grad_out.sum(0)
    dalpha = (g[i] * h[j]).sum(-1)                                   # [E',H]
    D = torch.zeros(M, H, dtype=x.dtype).index_add_(0, i, alpha * dalpha)
    da = alpha * (dalpha - D[i])
    dz = da * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, NEG_SLOPE))
    ds_dst = torch.zeros(M, H, dtype=x.dtype).index_add_(0, i, dz)
    ds_src = torch.zeros(M, H, dtype=x.dtype).index_add_(0, j, dz)
    dh = torch.zeros(M, H, C, dtype=x.dtype).index_add_(0, j, alpha.unsqueeze(-1) * g[i])
    dh = dh + ds_src.unsqueeze(-1) * att_src + ds_dst.unsqueeze(-1) * att_dst
    datt_src = (ds_src.unsqueeze(-1) * h).sum(0, keepdim=True)
    datt_dst = (ds_dst.unsqueeze(-1) * h).sum(0, keepdim=True)
    dh2 = dh.reshape(M, H * C)
    return dict(dx=dh2 @ weight, dW=dh2.t() @ x, datt_src=datt_src, datt_dst=datt_dst, dbias=dbias)


# --------------------------------------------------------------------------- #
# parameter init and the model
# --------------------------------------------------------------------------- #
def _glorot_(t: Tensor) -> Tensor:
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    return t.uniform_(-bound, bound)


def _fan_in_uniform_(t: Tensor, fan_in: int) -> Tensor:
    bound = 1.0 / math.sqrt(fan_in)
    return t.uniform_(-bound, bound)


class _Proj(nn.Module):
    """bias-free projection, named like PyG's `lin_src` so keys match."""
    def __init__(self, fin: int, fout: int):
        super().__init__()
        self.weight = nn.Parameter(_glorot_(torch.empty(fout, fin)))


class OracleGATConv(nn.Module):
    def __init__(self, fin: int, C: int, heads: int, concat: bool):
        super().__init__()
        self.heads, self.C, self.concat = heads, C, concat
        self.lin_src = _Proj(fin, heads * C)
        self.lin_dst = self.lin_src                      # PyG 2.3/2.4 aliasing
        self.att_src = nn.Parameter(_glorot_(torch.empty(1, heads, C)))
        self.att_dst = nn.Parameter(_glorot_(torch.empty(1, heads, C)))
        self.bias = nn.Parameter(torch.zeros(heads * C if concat else C))

    def forward(self, x, edge_index, edge_attr=None):    # edge_attr carried and ignored (no lin_edge)
        return gat_conv(x, edge_index, self.lin_src.weight, self.att_src, self.att_dst, self.bias,
                        self.heads, self.concat)


class _Affine(nn.Module):
    def __init__(self, fin: int, fout: int):
        super().__init__()
        self.weight = nn.Parameter(_fan_in_uniform_(torch.empty(fout, fin), fin))
        self.bias = nn.Parameter(_fan_in_uniform_(torch.empty(fout), fin))

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class OracleGResBlockMeanConv(nn.Module):
    """GraphModels.py:454-468."""
    def __init__(self, in_dim: int, out_dim: int, hc: int):
        super().__init__()
        self.conv1 = OracleGATConv(in_dim, hc, 2, concat=True)
        self.conv2 = OracleGATConv(hc * 2, out_dim, 1, concat=False)

    def forward(self, x, edge_index, edge_attr=None):
        x0 = x.clone()
        x = self.conv1(x, edge_index, edge_attr).relu()
        x = self.conv2(x, edge_index, edge_attr)
        x = simple_conv_mean(x, edge_index) + x0
        return F.relu(x)


class GATResOracle(nn.Module):
    """GraphModels.py:471-494 with PyG-2.3-style state_dict keys."""
    def __init__(self, name: str = "GATResMeanConv", num_blocks: int = 5, nc: int = 32):
        super().__init__()
        self.name, self.num_blocks, self.nc = name, num_blocks, nc
        self.lin0 = _Affine(1, nc)
        self.blocks = nn.ModuleList(OracleGResBlockMeanConv(nc, nc, nc) for _ in range(num_blocks))
        self.lin1 = _Affine(nc, 1)

    def forward(self, x, edge_index, batch=None, edge_attr=None):
        x = self.lin0(x)
        for blk in self.blocks:
            x = blk(x, edge_index, edge_attr)
        return self.lin1(x)                               # no output activation (:493 commented out)


class GATOracle(nn.Module):
    """The reference's plain GAT baseline, GraphModels.py:210-230 (`config_gat`, ConfigModels.py:96-103): GATConv layers
    only, two heads of nc channels, no activation in between, last layer one head of out_channels."""
    def __init__(self, name: str = "GAT", num_blocks: int = 10, nc: int = 32, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_blocks, self.name = num_blocks, f"{name}_{num_blocks}b_{nc}c"
        blocks = []
        for i in range(num_blocks):
            if i == 0:
                blocks.append(OracleGATConv(in_channels, nc, 2, True))
            elif i == num_blocks - 1:
                blocks.append(OracleGATConv(2 * nc, out_channels, 1, True))
            else:
                blocks.append(OracleGATConv(2 * nc, nc, 2, True))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x, edge_index, batch=None, edge_attr=None):
        for blk in self.blocks:
            x = blk(x, edge_index)
        return x


class OracleGResBlockConv(nn.Module):
    """GraphModels.py:548-561: conv1 -> ReLU -> conv2 -> + x_0 -> ReLU (no mean convolution)."""
    def __init__(self, in_dim: int, out_dim: int, hc: int):
        super().__init__()
        self.conv1 = OracleGATConv(in_dim, hc, 2, concat=True)
        self.conv2 = OracleGATConv(hc * 2, out_dim, 1, concat=False)

    def forward(self, x, edge_index, edge_attr=None):
        x0 = x.clone()
        x = self.conv1(x, edge_index, edge_attr).relu()
        x = self.conv2(x, edge_index, edge_attr)
        return F.relu(x + x0)


class GATConvNetOracle(nn.Module):
    """GraphModels.py:15-46: GATConv layers with Linear skips, ReLU + dropout(0.5) between them, sigmoid at the end.
    `dropout_masks` (keep masks, one per hidden layer) stands in for the random draw so training mode is comparable."""
    def __init__(self, net_params: dict):
        super().__init__()
        self.net_params = net_params
        heads, hid = net_params["heads"], net_params["hidden_dim"]
        self.convs = nn.ModuleList()
        fin = net_params["input_dim"]
        for _ in range(net_params["num_layers"] - 1):
            self.convs.append(OracleGATConv(fin, hid, heads, True))
            fin = heads * hid
        self.convs.append(OracleGATConv(heads * hid, net_params["out_dim"], 1, False))
        self.skips = nn.ModuleList([_Affine(net_params["input_dim"], heads * hid)])
        for _ in range(net_params["num_layers"] - 2):
            self.skips.append(_Affine(heads * hid, heads * hid))
        self.skips.append(_Affine(heads * hid, net_params["out_dim"]))

    def forward(self, x, edge_index, batch=None, dropout_masks=None):
        for i in range(self.net_params["num_layers"] - 1):
            x = F.relu(self.convs[i](x, edge_index) + self.skips[i](x))
            if dropout_masks is not None:
                x = x * dropout_masks[i] * 2.0
            else:
                x = F.dropout(x, p=0.5, training=self.training)
        x = self.convs[-1](x, edge_index) + self.skips[-1](x)
        return torch.sigmoid(x)


def make_gat_oracle(num_blocks: int = 10, nc: int = 32, seed: int = 0) -> GATOracle:
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = GATOracle(num_blocks=num_blocks, nc=nc)
    with torch.no_grad():
        for blk in m.blocks:
            blk.bias.uniform_(-0.1, 0.1)
    torch.random.set_rng_state(g)
    return m


def make_oracle(num_blocks: int, nc: int, seed: int = 0, dtype=torch.float32) -> GATResOracle:
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = GATResOracle(num_blocks=num_blocks, nc=nc)
    # biases of GATConv are zero at init; perturb them a little so parity tests
    # exercise the bias path (deterministic under the same seed).
    with torch.no_grad():
        for blk in m.blocks:
            blk.conv1.bias.uniform_(-0.1, 0.1)
            blk.conv2.bias.uniform_(-0.1, 0.1)
    torch.random.set_rng_state(g)
    return m.to(dtype)


# --------------------------------------------------------------------------- #
# caller semantics: collation, mask, loss  (train.py:159-185, auxil.py:143-182)
# --------------------------------------------------------------------------- #
def collate_edge_index(template_ei: Tensor, num_nodes: int, batch: int) -> Tensor:
    """PyG Batch: edge_index = cat(ei + b*N) along dim 1."""
    return torch.cat([template_ei + b * num_nodes for b in range(batch)], dim=1)


def mask_nodes(num_nodes: int, mask_rate: float, rng: np.random.RandomState) -> np.ndarray:
    """Exact-count random mask, one snapshot (auxil.py:143-163 with required_idx=[])."""
    k = int(num_nodes * mask_rate)
    assert k > 0
    idx = rng.choice(num_nodes, k, replace=False)
    m = np.zeros(num_nodes, dtype=bool)
    m[idx] = True
    return m


def generate_batch_mask(num_nodes: int, batch: int, mask_rate: float, seed: int = 1234) -> np.ndarray:
    rng = np.random.RandomState(seed)
    return np.hstack([mask_nodes(num_nodes, mask_rate, rng) for _ in range(batch)])


def synthetic_snapshots(num_nodes: int, batch: int, mask_rate: float = 0.95, seed: int = 1234,
                        dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """y ~ N(0,1) [B*N,1]; x = y with masked nodes zeroed; mask bool [B*N]."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(batch * num_nodes, 1, generator=g, dtype=torch.float32).to(dtype)
    mask = torch.from_numpy(generate_batch_mask(num_nodes, batch, mask_rate, seed))
    x = y.clone()
    x[mask] = 0
    return x, y, mask


def train_step_loss_and_grads(model: nn.Module, x: Tensor, y: Tensor, mask: Tensor, edge_index: Tensor
                              ) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """forward + MSE over masked nodes + backward (train.py:175-185)."""
    model.zero_grad(set_to_none=True)
    out = model(x, edge_index, None, None)
    loss = F.mse_loss(out[mask], y[mask])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return out.detach(), loss.detach(), grads


def adam_reference_step(params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, wd=6e-6,
                        beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (L2 weight decay, no amsgrad) on flat tensors; train.py:348."""
    g = grads + wd * params
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    params.addcdiv_(exp_avg, denom, value=-lr / bc1)
    return params
