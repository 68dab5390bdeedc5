"""Drop-in ``GraphModels`` surface for the GATRes hot path.

Same public names, constructor signatures, ``forward(x, edge_index, batch=None,
edge_attr=None)`` and PyG-compatible ``state_dict`` keys as
/root/reference/gnn_pressure_estimation/GraphModels.py:454-494, but no
``torch_geometric`` import: the arithmetic runs in the sm_100a kernels behind
``torch.ops.gatres`` (CUDA only — a CPU tensor raises).

    GATResMeanConv.forward  -> one fused-stack op (gatres_forward / gatres_backward)
    GResBlockMeanConv, GATConv, SimpleConv, Linear
                            -> the same kernels op by op (used by the parity tests
                               and by callers that compose blocks themselves)
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn.functional as F  # noqa: F401  (kept for callers that do `from GraphModels import *`)
from torch import Tensor, nn

from . import ops as _gops
from .graph import Topology, TopologyCache

__all__ = ["GATResMeanConv", "GResBlockMeanConv", "GResBlockConv", "GAT", "GATConvNet", "GATConv", "SimpleConv",
           "Linear"]

_SHARED_TOPOLOGIES = TopologyCache()


def _require_cuda(x: Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: this is the B200 build of the GATRes hot path — CUDA tensors only, "
                           "there is no CPU / PyG fallback (move the model and the batch to the GPU)")


def _glorot_(t: Tensor) -> Tensor:
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)


class Linear(nn.Module):
    """``torch_geometric.nn.dense.linear.Linear`` stand-in: kaiming-uniform(a=sqrt 5)
    weight and U(+-1/sqrt(fan_in)) bias by default, glorot for GATConv's projection."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, weight_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.weight_initializer = weight_initializer
        self.reset_parameters()

    def reset_parameters(self) -> None:
        if self.weight_initializer == "glorot":
            _glorot_(self.weight)
        else:
            b = 1.0 / math.sqrt(self.in_channels)
            with torch.no_grad():
                self.weight.uniform_(-b, b)
        if self.bias is not None:
            b = 1.0 / math.sqrt(self.in_channels)
            with torch.no_grad():
                self.bias.uniform_(-b, b)

    def forward(self, x: Tensor) -> Tensor:
        # lin0 / lin1 of GATRes (GraphModels.py:477,484) run on the library's encoder / decoder kernels when a model
        # is composed module by module; other widths (the skip connections of the sibling GATConvNet) are plain
        # dense layers off the GATRes path and use torch's own CUDA GEMM.
        _require_cuda(x, "Linear")
        if self.bias is not None and x.dim() == 2 and x.dtype == torch.float32:
            if self.in_channels == 1 and self.out_channels in (32, 64, 128):
                return _gops.encoder(x, self.weight, self.bias)
            if self.out_channels == 1 and self.in_channels in (32, 64, 128):
                return _gops.decoder(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class GATConv(nn.Module):
    """``GATConv(in, out, heads, concat)`` with PyG's defaults (negative_slope 0.2,
    add_self_loops, bias, no dropout, no edge features).  Parameters are named as
    PyG 2.3/2.4 names them (``lin_src`` with ``lin_dst`` aliasing it); PyG >= 2.5
    checkpoints (``lin.weight``) load too."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.lin_src.reset_parameters()
        _glorot_(self.att_src)
        _glorot_(self.att_dst)
        with torch.no_grad():
            self.bias.zero_()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        new = prefix + "lin.weight"                      # PyG >= 2.5 naming
        if new in state_dict:
            w = state_dict.pop(new)
            state_dict.setdefault(prefix + "lin_src.weight", w)
            state_dict.setdefault(prefix + "lin_dst.weight", w)
        elif prefix + "lin_src.weight" in state_dict:
            state_dict.setdefault(prefix + "lin_dst.weight", state_dict[prefix + "lin_src.weight"])
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None, *, _relu: bool = False
                ) -> Tensor:
        # edge_attr is accepted and ignored exactly as in the reference (no lin_edge)
        _require_cuda(x, "GATConv")
        topo, B = _SHARED_TOPOLOGIES.resolve(x.size(0), edge_index)
        return _gops.gat_conv(x, self.lin_src.weight, self.att_src, self.att_dst, self.bias, topo, B, self.heads,
                              self.concat, relu=_relu)


class SimpleConv(nn.Module):
    """``SimpleConv(aggr="mean")``; the residual add + ReLU that always follow it
    in GATRes are fused in via ``forward_residual_relu``."""

    def __init__(self, aggr: str = "mean"):
        super().__init__()
        if aggr != "mean":
            raise NotImplementedError("only aggr='mean' is on the GATRes path")
        self.aggr = aggr

    def forward_residual_relu(self, x: Tensor, edge_index: Tensor, x0: Tensor) -> Tensor:
        _require_cuda(x, "SimpleConv")
        topo, B = _SHARED_TOPOLOGIES.resolve(x.size(0), edge_index)
        return _gops.mean_res(x, x0, topo, B)


class GResBlockMeanConv(nn.Module):
    """GraphModels.py:454-468 of the reference."""

    def __init__(self, in_dim: int, out_dim: int, hc: int):
        super().__init__()
        self.conv1 = GATConv(in_dim, hc, 2, concat=True)
        self.conv2 = GATConv(hc * 2, out_dim, 1, concat=False)
        self.mean_conv = SimpleConv(aggr="mean")

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tensor:
        x_0 = x
        x = self.conv1(x, edge_index, edge_attr, _relu=True)
        x = self.conv2(x, edge_index, edge_attr)
        return self.mean_conv.forward_residual_relu(x, edge_index, x_0)


class GResBlockConv(nn.Module):
    """GraphModels.py:548-561 of the reference: the residual block without the mean convolution
    (conv1 -> ReLU -> conv2 -> + x_0 -> ReLU); same GATConv kernels as `GResBlockMeanConv`."""

    def __init__(self, in_dim: int, out_dim: int, hc: int):
        super().__init__()
        self.conv1 = GATConv(in_dim, hc, 2, concat=True)
        self.conv2 = GATConv(hc * 2, out_dim, 1, concat=False)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tensor:
        x_0 = x
        x = self.conv1(x, edge_index, edge_attr, _relu=True)
        x = self.conv2(x, edge_index, edge_attr)
        return F.relu(x + x_0)


class GATConvNet(nn.Module):
    """GraphModels.py:15-46 of the reference: `num_layers` GATConv layers (heads x hidden_dim, concat; the last one a
    single head of out_dim, mean) each with a Linear skip connection, ReLU + dropout(0.5) between layers, sigmoid at
    the end.  The GATConv layers run on the fused kernels (widths zero-padded to built shapes, see `ops.gat_conv`);
    skips, dropout and the sigmoid are elementwise / dense torch ops — this model is a baseline next to the hot path.
    `dropout_masks` (one [M, heads*hidden_dim] 0/1 tensor per hidden layer) replaces the random draw for parity tests."""

    def __init__(self, net_params: dict):
        super().__init__()
        self.net_params = net_params
        heads, hid = net_params["heads"], net_params["hidden_dim"]
        self.convs = nn.ModuleList()
        in_channels = net_params["input_dim"]
        for _ in range(net_params["num_layers"] - 1):
            self.convs.append(GATConv(in_channels, hid, heads=heads, concat=True))
            in_channels = heads * hid
        self.convs.append(GATConv(heads * hid, net_params["out_dim"], heads=1, concat=False))
        self.skips = nn.ModuleList()
        self.skips.append(Linear(net_params["input_dim"], heads * hid))
        for _ in range(net_params["num_layers"] - 2):
            self.skips.append(Linear(heads * hid, heads * hid))
        self.skips.append(Linear(heads * hid, net_params["out_dim"]))

    def forward(self, x: Tensor, edge_index: Tensor, batch: Optional[Tensor] = None,
                dropout_masks: Optional[List[Tensor]] = None) -> Tensor:
        for i in range(self.net_params["num_layers"] - 1):
            x = F.relu(self.convs[i](x, edge_index) + self.skips[i](x))
            if dropout_masks is not None:
                x = x * dropout_masks[i] * 2.0                       # dropout(p=0.5) with a given keep mask
            else:
                x = F.dropout(x, p=0.5, training=self.training)
        x = self.convs[-1](x, edge_index) + self.skips[-1](x)
        return torch.sigmoid(x)


class GAT(nn.Module):
    """The reference's plain GAT baseline (/root/reference/gnn_pressure_estimation/GraphModels.py:210-230,
    `config_gat`, ConfigModels.py:96-103): `num_blocks` GATConv layers, two heads of `nc` channels, no activation
    between them; first layer from `in_channels`, last layer one head of `out_channels`.  Runs on the same fused
    kernels as GATRes; the 1-wide first / last layers are zero-padded to built kernel shapes."""

    def __init__(self, name: str = "GAT", num_blocks: int = 10, nc: int = 32, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_blocks = num_blocks
        self.name = f"{name}_{num_blocks}b_{nc}c"
        blocks = []
        for i in range(num_blocks):
            if i == 0:
                blocks.append(GATConv(in_channels, nc, heads=2, concat=True))
            elif i == num_blocks - 1:
                blocks.append(GATConv(2 * nc, out_channels, heads=1, concat=True))
            else:
                blocks.append(GATConv(2 * nc, nc, heads=2, concat=True))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x: Tensor, edge_index: Tensor, batch: Optional[Tensor] = None, edge_attr: Optional[Tensor] = None) -> Tensor:
        for blk in self.blocks:
            x = blk(x, edge_index)
        return x


class GATResMeanConv(nn.Module):
    """GraphModels.py:471-494 of the reference; forward runs as one fused-stack op."""

    def __init__(self, name: str = "GATResMeanConv", num_blocks: int = 5, nc: int = 32):
        super().__init__()
        self.num_blocks = num_blocks
        self.nc = nc
        self.lin0 = Linear(1, nc)
        self.blocks = nn.ModuleList()
        self.name = name
        for _ in range(self.num_blocks):
            self.blocks.append(GResBlockMeanConv(nc, nc, nc))
        self.lin1 = Linear(nc, 1)
        self._flat: Optional[Tensor] = None
        self._topologies = TopologyCache()
        # True: parameter gradients are reduced in a fixed order (bitwise reproducible, slower backward);
        # False: atomic accumulation, like the reference's scatter-add backward on GPU.
        self.deterministic = False

    # -- flat parameter storage (layout documented in include/gatres_b200.h) ------------
    def ordered_parameters(self) -> List[nn.Parameter]:
        ps: List[nn.Parameter] = [self.lin0.weight, self.lin0.bias]
        for b in self.blocks:
            ps += [b.conv1.lin_src.weight, b.conv1.att_src, b.conv1.att_dst, b.conv1.bias,
                   b.conv2.lin_src.weight, b.conv2.att_src, b.conv2.att_dst, b.conv2.bias]
        return ps + [self.lin1.weight, self.lin1.bias]

    def flat_parameters(self) -> Tensor:
        """One contiguous fp32 buffer all parameters are views of (re-packed if
        ``.to()`` / an optimizer swap broke the aliasing)."""
        ps = self.ordered_parameters()
        flat = self._flat
        ok = flat is not None and flat.device == ps[0].device
        if ok:
            base, off = flat.data_ptr(), 0
            for p in ps:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            if any(p.dtype != torch.float32 for p in ps):
                raise RuntimeError("GATResMeanConv (B200 build) computes in fp32; got parameters of another dtype")
            with torch.no_grad():
                flat = torch.cat([p.detach().reshape(-1) for p in ps]).contiguous()
                off = 0
                for p in ps:
                    p.data = flat[off:off + p.numel()].view(p.shape)
                    off += p.numel()
            assert flat.numel() == _gops.param_count(self.num_blocks, self.nc)
            self._flat = flat
        return flat

    def _forward_composed(self, x: Tensor, edge_index: Tensor) -> Tensor:
        """GraphModels.py:486-494 literally: lin0 -> blocks -> lin1, every module on its own kernels."""
        x = self.lin0(x)
        for blk in self.blocks:
            x = blk(x, edge_index, None)
        return self.lin1(x)

    def set_topology(self, edge_index: Tensor, num_nodes: int) -> Topology:
        """Optional: register the template graph up front (otherwise it is inferred
        from the first collated batch)."""
        return self._topologies.set_template(edge_index, num_nodes)

    def forward(self, x: Tensor, edge_index: Tensor, batch: Optional[Tensor] = None,
                edge_attr: Optional[Tensor] = None) -> Tensor:
        _require_cuda(x, "GATResMeanConv")
        if x.dim() != 2 or x.size(1) != 1:
            raise ValueError(f"GATResMeanConv expects x of shape [num_nodes, 1], got {tuple(x.shape)}")
        topo, B = self._topologies.resolve(x.size(0), edge_index, batch)
        if not topo.shares_one_csr:
            # a template with self loops: GATConv rewrites them, SimpleConv(mean) keeps them, so the
            # one-CSR fused stack does not apply — run the same kernels module by module (two CSR views)
            return self._forward_composed(x, edge_index)
        flat = self.flat_parameters()
        out = _gops.gatres_model(x.reshape(-1), flat, self.ordered_parameters(), topo, B, self.num_blocks, self.nc,
                                 poison=self._topologies.mismatch, deterministic=self.deterministic)
        return out.view(-1, 1)
