"""``torch.ops.gatres.*`` — the thin operator layer over the C ABI.

Every op is registered for the CUDA dispatch key only (``torch.library``), so a
CPU tensor fails loudly in the dispatcher: there is no CPU path, no PyG scatter
fallback and no Triton.  Each implementation allocates its outputs with torch,
takes raw pointers and the current stream, and makes exactly one C call.

The autograd-aware entry points the model uses are at the bottom
(``gat_conv``, ``mean_res``, ``gatres_model``); they mirror the operator
interface the reference gets from PyG
(/root/reference/gnn_pressure_estimation/GraphModels.py:458-466, :486-494).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import ModelDesc, call, ptr, stream

_L = torch.library.Library("gatres", "DEF")


def _def(schema: str, fn) -> None:
    name = schema.split("(", 1)[0]
    _L.define(schema)
    _L.impl(name, fn, "CUDA")


def _f32(t: Tensor, what: str) -> Tensor:
    if t.dtype != torch.float32:
        raise _lib.GatresError(f"{what}: the sm_100a kernels compute in fp32, got {t.dtype}")
    return t.contiguous()


def a4(n: int) -> int:
    return (n + 3) & ~3


_SM = None


def sm_count() -> int:
    global _SM
    if _SM is None:
        _SM = int(_lib.load().gatres_sm_count())
    return _SM


def grad_slots(rows: int) -> int:
    """CTAs used by gradient-producing kernels (rows of the `partial` buffer)."""
    return max(1, min(sm_count(), (rows + 63) // 64))


# ----------------------------------------------------------------------------
# graph
# ----------------------------------------------------------------------------
def _csr_build(edge_index: Tensor, N: int, keep_self_loops: bool = False) -> List[Tensor]:
    """keep_self_loops=False: GATConv's view (gatres_csr_build); True: SimpleConv's view (gatres_csr_build_mean)"""
    ei = edge_index.contiguous()
    if ei.dtype != torch.int64 or ei.dim() != 2 or ei.size(0) != 2:
        raise _lib.GatresError("edge_index must be int64 [2, E]")
    E = ei.size(1)
    dev = ei.device
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    rowptr_t = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E + N, dtype=torch.int32, device=dev)
    col_t = torch.empty(E + N, dtype=torch.int32, device=dev)
    info = torch.empty(4, dtype=torch.int32, device=dev)
    nbytes = int(_lib.load().gatres_csr_scratch_bytes(E, N))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    call("gatres_csr_build_mean" if keep_self_loops else "gatres_csr_build", ptr(ei), E, N, ptr(rowptr), ptr(col),
         ptr(rowptr_t), ptr(col_t), ptr(info), ptr(scratch), nbytes, stream())
    return [rowptr, col, rowptr_t, col_t, info]


_def("csr_build(Tensor edge_index, int N, bool keep_self_loops=False) -> Tensor[]", _csr_build)


def _check_replicated(ei_batch: Tensor, ei_tmpl: Tensor, B: int, N: int, mismatch: Tensor) -> None:
    E = ei_tmpl.size(1)
    if ei_batch.size(1) != B * E:
        raise _lib.GatresError("check_replicated: edge_index has the wrong number of columns")
    call("gatres_check_replicated", ptr(ei_batch.contiguous()), ptr(ei_tmpl.contiguous()), B, E, N, ptr(mismatch),
         stream())


_def("check_replicated(Tensor ei_batch, Tensor ei_tmpl, int B, int N, Tensor(a!) mismatch) -> ()", _check_replicated)


# ----------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------
def _linear_att_fwd(x: Tensor, W: Tensor, att_src: Tensor, att_dst: Tensor, H: int, C_: int) -> List[Tensor]:
    x, W = _f32(x, "linear_att_fwd"), _f32(W, "linear_att_fwd")
    M, K = x.shape
    h = torch.empty(M, H * C_, dtype=torch.float32, device=x.device)
    s_src = torch.empty(M, H, dtype=torch.float32, device=x.device)
    s_dst = torch.empty(M, H, dtype=torch.float32, device=x.device)
    call("gatres_linear_att_fwd", ptr(x), ptr(W), ptr(_f32(att_src, "att")), ptr(_f32(att_dst, "att")), ptr(h),
         ptr(s_src), ptr(s_dst), M, K, H, C_, stream())
    return [h, s_src, s_dst]


_def("linear_att_fwd(Tensor x, Tensor W, Tensor att_src, Tensor att_dst, int H, int C) -> Tensor[]", _linear_att_fwd)


def _gat_agg_fwd(rowptr: Tensor, col: Tensor, h: Tensor, s_src: Tensor, s_dst: Tensor, bias: Tensor, B: int, N: int,
                 H: int, C_: int, relu: bool, save_stats: bool) -> List[Tensor]:
    h = _f32(h, "gat_agg_fwd")
    M = B * N
    out = torch.empty(M, H * C_, dtype=torch.float32, device=h.device)
    m = torch.empty(M, H, dtype=torch.float32, device=h.device) if save_stats else None
    l = torch.empty(M, H, dtype=torch.float32, device=h.device) if save_stats else None
    call("gatres_gat_agg_fwd", ptr(rowptr), ptr(col), ptr(h), ptr(s_src.contiguous()), ptr(s_dst.contiguous()),
         ptr(_f32(bias, "bias")), ptr(out), ptr(m), ptr(l), B, N, col.numel(), H, C_, int(relu), stream())
    empty = out.new_empty(0)
    return [out, m if save_stats else empty, l if save_stats else empty]


_def("gat_agg_fwd(Tensor rowptr, Tensor col, Tensor h, Tensor s_src, Tensor s_dst, Tensor bias, int B, int N, "
     "int H, int C, bool relu, bool save_stats) -> Tensor[]", _gat_agg_fwd)


def _gat_agg_bwd(rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor, g: Tensor, h: Tensor, s_src: Tensor,
                 s_dst: Tensor, m: Tensor, l: Tensor, att_src: Tensor, att_dst: Tensor, B: int, N: int, H: int,
                 C_: int) -> List[Tensor]:
    g, h = _f32(g, "gat_agg_bwd"), _f32(h, "gat_agg_bwd")
    M, F = B * N, H * C_
    dev = h.device
    P = 3 * F                                   # [datt_src | datt_dst | dbias], multiple of 4
    grads = torch.zeros(P, dtype=torch.float32, device=dev)      # atomic accumulation target (slots = 0)
    rec = torch.empty(M, H, 4, dtype=torch.float32, device=dev)
    ds_dst = torch.empty(M, H, dtype=torch.float32, device=dev)
    dh = torch.empty(M, F, dtype=torch.float32, device=dev)
    call("gatres_gat_agg_bwd", ptr(rowptr), ptr(col), ptr(rowptr_t), ptr(col_t), ptr(g), ptr(h),
         ptr(s_src.contiguous()), ptr(s_dst.contiguous()), ptr(m.contiguous()), ptr(l.contiguous()),
         ptr(_f32(att_src, "att")), ptr(_f32(att_dst, "att")), ptr(rec), ptr(ds_dst), ptr(dh), ptr(grads),
         P, 0, 0, F, 2 * F, B, N, col.numel(), H, C_, stream())
    return [dh, grads[:F].view(1, H, C_), grads[F:2 * F].view(1, H, C_), grads[2 * F:]]


_def("gat_agg_bwd(Tensor rowptr, Tensor col, Tensor rowptr_t, Tensor col_t, Tensor g, Tensor h, Tensor s_src, "
     "Tensor s_dst, Tensor m, Tensor l, Tensor att_src, Tensor att_dst, int B, int N, int H, int C) -> Tensor[]",
     _gat_agg_bwd)


def _linear_bwd(dh: Tensor, x: Tensor, W: Tensor, add: Optional[Tensor], relu_ref: Optional[Tensor], H: int,
                C_: int) -> List[Tensor]:
    dh, x, W = _f32(dh, "linear_bwd"), _f32(x, "linear_bwd"), _f32(W, "linear_bwd")
    M, K = x.shape
    NO = H * C_
    dev = x.device
    S = grad_slots(M)
    P = NO * K
    partial = torch.empty(S, P, dtype=torch.float32, device=dev)
    dx = torch.empty(M, K, dtype=torch.float32, device=dev)
    call("gatres_linear_bwd", ptr(dh), ptr(x), ptr(W), ptr(add), ptr(relu_ref), ptr(dx), ptr(partial), P, S, 0, M, K,
         H, C_, stream())
    dW = torch.empty(P, dtype=torch.float32, device=dev)
    call("gatres_reduce_partials", ptr(partial), P, S, 0, P, ptr(dW), stream())
    return [dx, dW.view(NO, K)]


_def("linear_bwd(Tensor dh, Tensor x, Tensor W, Tensor? add, Tensor? relu_ref, int H, int C) -> Tensor[]", _linear_bwd)


def _mean_res_fwd(rowptr: Tensor, col: Tensor, z: Tensor, x0: Tensor, B: int, N: int) -> Tensor:
    z, x0 = _f32(z, "mean_res_fwd"), _f32(x0, "mean_res_fwd")
    out = torch.empty_like(z)
    call("gatres_mean_res_fwd", ptr(rowptr), ptr(col), ptr(z), ptr(x0), ptr(out), B, N, z.size(1), stream())
    return out


_def("mean_res_fwd(Tensor rowptr, Tensor col, Tensor z, Tensor x0, int B, int N) -> Tensor", _mean_res_fwd)


def _mean_res_bwd(rowptr: Tensor, rowptr_t: Tensor, col_t: Tensor, g_out: Tensor, out: Tensor, B: int, N: int
                  ) -> List[Tensor]:
    g_out, out = _f32(g_out, "mean_res_bwd"), _f32(out, "mean_res_bwd")
    dz = torch.empty_like(out)
    dres = torch.empty_like(out)
    call("gatres_mean_res_bwd", ptr(rowptr), ptr(rowptr_t), ptr(col_t), ptr(g_out), ptr(out), ptr(dz), ptr(dres), B, N,
         out.size(1), stream())
    return [dz, dres]


_def("mean_res_bwd(Tensor rowptr, Tensor rowptr_t, Tensor col_t, Tensor g_out, Tensor out, int B, int N) -> Tensor[]",
     _mean_res_bwd)


# ----------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------
def _desc(num_blocks: int, nc: int, N: int, B: int, rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor,
          poison: Optional[Tensor], deterministic: bool = False, plan: Optional[List[Tensor]] = None) -> ModelDesc:
    """slots > 0: parameter gradients via per-CTA partial rows + a fixed-order reduction (bitwise
    reproducible, small grids); slots = 0: atomic accumulation into the gradient buffer (full grids).
    plan: [perm, p_rowptr, p_col, p_rowptr_t, p_col_t, ecap(int32[4], host)] of graph.LocalityPlan, or None."""
    d = ModelDesc(num_blocks, nc, N, grad_slots(B * N) if deterministic else 0, col.numel(), 0, B, ptr(rowptr), ptr(col),
                  ptr(rowptr_t), ptr(col_t), ptr(poison))
    if plan:
        d.perm, d.p_rowptr, d.p_col, d.p_rowptr_t, d.p_col_t = (ptr(t) for t in plan[:5])
        for q, v in enumerate(plan[5].tolist()):
            d.p_ecap[q] = int(v)
    return d


def plan_tensors(topo) -> List[Tensor]:
    """graph.LocalityPlan of a Topology as the tensor list the model ops take ([] = no plan)"""
    p = topo.plan
    if p is None:
        return []
    return [p.perm, p.rowptr, p.col, p.rowptr_t, p.col_t, torch.tensor(p.ecap, dtype=torch.int32)]


def param_count(num_blocks: int, nc: int) -> int:
    return int(_lib.load().gatres_param_count(num_blocks, nc))


def _model_forward(params: Tensor, x: Tensor, rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor,
                   poison: Optional[Tensor], num_blocks: int, nc: int, N: int, B: int, training: bool,
                   deterministic: bool, plan: List[Tensor]) -> List[Tensor]:
    params, x = _f32(params, "model_forward"), _f32(x, "model_forward")
    M = B * N
    if x.numel() != M:
        raise _lib.GatresError(f"model_forward: x has {x.numel()} rows, expected B*N = {M}")
    if params.numel() != param_count(num_blocks, nc):
        raise _lib.GatresError("model_forward: flat parameter buffer has the wrong size")
    lib = _lib.load()
    # the same descriptor (gradient mode, locality plan) as the backward: both sides must pick the same kernels
    d = _desc(num_blocks, nc, N, B, rowptr, col, rowptr_t, col_t, poison, deterministic, plan)
    dev = x.device
    out = torch.empty(M, dtype=torch.float32, device=dev)
    saved = torch.empty(int(lib.gatres_saved_floats(C.byref(d))) if training else 0, dtype=torch.float32, device=dev)
    n_scr = M * nc if training else int(lib.gatres_scratch_floats(C.byref(d), 0))
    scratch = torch.empty(n_scr, dtype=torch.float32, device=dev)
    call("gatres_forward", C.byref(d), ptr(params), ptr(x), ptr(out), ptr(saved) if training else None, ptr(scratch),
         stream())
    return [out, saved]


_def("model_forward(Tensor params, Tensor x, Tensor rowptr, Tensor col, Tensor rowptr_t, Tensor col_t, "
     "Tensor? poison, int num_blocks, int nc, int N, int B, bool training, bool deterministic, Tensor[] plan) -> Tensor[]",
     _model_forward)


def _model_backward(params: Tensor, x: Tensor, saved: Tensor, d_out: Tensor, rowptr: Tensor, col: Tensor,
                    rowptr_t: Tensor, col_t: Tensor, num_blocks: int, nc: int, N: int, B: int,
                    deterministic: bool, plan: List[Tensor]) -> Tensor:
    params, x, d_out = _f32(params, "model_backward"), _f32(x, "model_backward"), _f32(d_out, "model_backward")
    lib = _lib.load()
    d = _desc(num_blocks, nc, N, B, rowptr, col, rowptr_t, col_t, None, deterministic, plan)
    dev = x.device
    P = params.numel()
    grads = torch.empty(P, dtype=torch.float32, device=dev)
    partial = torch.empty(d.slots * a4(P), dtype=torch.float32, device=dev) if deterministic else None
    scratch = torch.empty(int(lib.gatres_scratch_floats(C.byref(d), 1)), dtype=torch.float32, device=dev)
    call("gatres_backward", C.byref(d), ptr(params), ptr(x), ptr(saved), ptr(d_out), ptr(partial), ptr(grads),
         ptr(scratch), stream())
    return grads


_def("model_backward(Tensor params, Tensor x, Tensor saved, Tensor d_out, Tensor rowptr, Tensor col, "
     "Tensor rowptr_t, Tensor col_t, int num_blocks, int nc, int N, int B, bool deterministic, Tensor[] plan) -> Tensor",
     _model_backward)


# ----------------------------------------------------------------------------
# caller-side fusions
# ----------------------------------------------------------------------------
def _masked_mse(out: Tensor, y: Tensor, mask: Tensor, count: int) -> List[Tensor]:
    out, y = _f32(out, "masked_mse"), _f32(y, "masked_mse")
    mask = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous()
    M = out.numel()
    d_out = torch.empty_like(out)
    loss = torch.empty(1, dtype=torch.float32, device=out.device)
    part = torch.empty(1024, dtype=torch.float32, device=out.device)
    call("gatres_masked_mse", ptr(out), ptr(y), ptr(mask), M, count, ptr(d_out), ptr(loss), ptr(part), stream())
    return [loss, d_out]


_def("masked_mse(Tensor out, Tensor y, Tensor mask, int count) -> Tensor[]", _masked_mse)


def _apply_mask(x: Tensor, mask: Tensor) -> Tensor:
    x = _f32(x, "apply_mask")
    mask = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous()
    out = torch.empty_like(x)
    call("gatres_apply_mask", ptr(x), ptr(mask), ptr(out), x.numel(), stream())
    return out


_def("apply_mask(Tensor x, Tensor mask) -> Tensor", _apply_mask)


def _adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: Tensor, lr: float,
               beta1: float, beta2: float, eps: float, weight_decay: float, grad_scale: float) -> None:
    call("gatres_adam_step", ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), ptr(step), params.numel(), lr,
         beta1, beta2, eps, weight_decay, grad_scale, stream())


_def("adam_step(Tensor(a!) params, Tensor grads, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, Tensor(d!) step, "
     "float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale) -> ()", _adam_step)

_ops = torch.ops.gatres


# ----------------------------------------------------------------------------
# autograd-aware entry points
# ----------------------------------------------------------------------------
class _GATConvFn(torch.autograd.Function):
    """GATConv.forward (projection + fused aggregation), SURVEY §A.2 / §A.4."""

    @staticmethod
    def forward(ctx, x, W, att_src, att_dst, bias, topo, B, H, C_, relu):
        h, s_src, s_dst = _ops.linear_att_fwd(x, W, att_src, att_dst, H, C_)
        need = any(ctx.needs_input_grad[:5])
        out, m, l = _ops.gat_agg_fwd(topo.rowptr, topo.col, h, s_src, s_dst, bias, B, topo.N, H, C_, relu, need)
        if need:
            ctx.save_for_backward(x, W, att_src, att_dst, h, s_src, s_dst, m, l, out)
            ctx.topo, ctx.B, ctx.H, ctx.C, ctx.relu = topo, B, H, C_, relu
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, W, att_src, att_dst, h, s_src, s_dst, m, l, out = ctx.saved_tensors
        t = ctx.topo
        g = g.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        dh, das, dad, db = _ops.gat_agg_bwd(t.rowptr, t.col, t.rowptr_t, t.col_t, g, h, s_src, s_dst, m, l,
                                            att_src, att_dst, ctx.B, t.N, ctx.H, ctx.C)
        dx, dW = _ops.linear_bwd(dh, x, W, None, None, ctx.H, ctx.C)
        return dx, dW, das, dad, db, None, None, None, None, None


# (heads, channels) -> input widths the projection kernels are built for (linear.cu dispatch tables)
_SUPPORTED_K = {(2, 32): (32, 64), (1, 32): (64,), (2, 64): (64,), (1, 64): (128,), (2, 128): (128,), (1, 128): (256,)}


def gat_conv(x: Tensor, W: Tensor, att_src: Tensor, att_dst: Tensor, bias: Tensor, topo, B: int, heads: int,
             concat: bool, relu: bool = False) -> Tensor:
    """GATConv on the fused kernels.  Layer shapes outside the kernels' tables (the 1-wide first / last layers of the
    reference's sibling `GAT` model, GraphModels.py:210-230) are zero-padded to the nearest built shape: padded input
    columns and padded output channels contribute exact zeros to the projection, the scores and the aggregation,
    and autograd slices their gradients away."""
    C_ = W.size(0) // heads
    if not concat and heads != 1:
        raise NotImplementedError("concat=False is implemented for heads=1 (the only use in GATRes)")
    Cp = next((c for c in (32, 64, 128) if c >= C_), None)
    if Cp is None or (heads, Cp) not in _SUPPORTED_K:
        raise NotImplementedError(f"GATConv with heads={heads}, out_channels={C_} has no kernel")
    K = x.size(1)
    Kp = next((k for k in _SUPPORTED_K[(heads, Cp)] if k >= K), None)
    if Kp is None:
        raise NotImplementedError(f"GATConv with in_channels={K}, heads={heads}, out_channels={C_} has no kernel")
    if Cp == C_ and Kp == K:
        return _GATConvFn.apply(x, W, att_src, att_dst, bias, topo, B, heads, C_, relu)
    pad = torch.nn.functional.pad
    Wp = pad(W.view(heads, C_, K), (0, Kp - K, 0, Cp - C_)).reshape(heads * Cp, Kp)
    asp, adp = pad(att_src, (0, Cp - C_)), pad(att_dst, (0, Cp - C_))
    bp = pad(bias.view(-1, C_), (0, Cp - C_)).reshape(-1)
    out = _GATConvFn.apply(pad(x, (0, Kp - K)), Wp, asp, adp, bp, topo, B, heads, Cp, relu)
    return out.view(out.size(0), -1, Cp)[:, :, :C_].reshape(out.size(0), -1)


class _EncoderFn(torch.autograd.Function):
    """PyG Linear(1, nc) — lin0 of GATRes (GraphModels.py:477,487): out[m, c] = x[m] w[c] + b[c]."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _f32(x, "encoder").reshape(-1)
        nc = w.numel()
        out = torch.empty(x.numel(), nc, dtype=torch.float32, device=x.device)
        call("gatres_encoder_fwd", ptr(x), ptr(_f32(w, "encoder").reshape(-1)), ptr(_f32(b, "encoder")), ptr(out), x.numel(), nc, stream())
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        nc = w.numel()
        g = _f32(g, "encoder_bwd")
        grads = torch.zeros(2 * nc, dtype=torch.float32, device=g.device)          # [dw | db], atomic accumulation
        call("gatres_encoder_bwd", ptr(g), ptr(x), ptr(grads), 2 * nc, 0, 0, nc, x.numel(), nc, stream())
        dx = (g @ w.reshape(nc, 1)) if ctx.needs_input_grad[0] else None           # model inputs never need it
        return dx, grads[:nc].view_as(w), grads[nc:]


class _DecoderFn(torch.autograd.Function):
    """PyG Linear(nc, 1) — lin1 of GATRes (GraphModels.py:484,492): out[m] = <x[m, :], w> + b."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _f32(x, "decoder")
        M, nc = x.shape
        out = torch.empty(M, dtype=torch.float32, device=x.device)
        call("gatres_decoder_fwd", ptr(x), ptr(_f32(w, "decoder").reshape(-1)), ptr(_f32(b, "decoder")), ptr(out), None, M, nc, stream())
        ctx.save_for_backward(x, w)
        return out.view(M, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        M, nc = x.shape
        g = _f32(g, "decoder_bwd").reshape(-1)
        P = a4(nc + 1)
        grads = torch.zeros(P, dtype=torch.float32, device=g.device)               # [dw | db]
        dx = torch.empty_like(x)
        call("gatres_decoder_bwd", ptr(g), ptr(x), ptr(w.reshape(-1).contiguous()), ptr(dx), ptr(grads), P, 0, 0, nc, M, nc, 0, stream())
        return dx, grads[:nc].view_as(w), grads[nc:nc + 1]


def encoder(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return _EncoderFn.apply(x, w, b)


def decoder(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return _DecoderFn.apply(x, w, b)


class _MeanResFn(torch.autograd.Function):
    """relu(SimpleConv(mean)(z) + x0), GraphModels.py:466-467."""

    @staticmethod
    def forward(ctx, z, x0, topo, B):
        rp, col, _, _ = topo.mean_view()
        out = _ops.mean_res_fwd(rp, col, z, x0, B, topo.N)
        ctx.save_for_backward(out)
        ctx.topo, ctx.B = topo, B
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        t = ctx.topo
        rp, _, rpt, colt = t.mean_view()
        dz, dres = _ops.mean_res_bwd(rp, rpt, colt, g.contiguous(), out, ctx.B, t.N)
        return dz, dres, None, None


def mean_res(z: Tensor, x0: Tensor, topo, B: int) -> Tensor:
    return _MeanResFn.apply(z, x0, topo, B)


class _ModelFn(torch.autograd.Function):
    """GATResMeanConv.forward as one op; backward returns every parameter
    gradient as a view into one flat buffer (same layout as the parameters)."""

    @staticmethod
    def forward(ctx, x, flat, topo, B, num_blocks, nc, poison, shapes, deterministic, *params):
        training = any(ctx.needs_input_grad[9:]) or ctx.needs_input_grad[1]
        plan = plan_tensors(topo)
        out, saved = _ops.model_forward(flat, x, topo.rowptr, topo.col, topo.rowptr_t, topo.col_t, poison, num_blocks,
                                        nc, topo.N, B, training, deterministic, plan)
        if training:
            ctx.save_for_backward(x, flat, saved)
            ctx.cfg = (topo, B, num_blocks, nc, shapes, deterministic, plan)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, flat, saved = ctx.saved_tensors
        topo, B, num_blocks, nc, shapes, deterministic, plan = ctx.cfg
        grads = _ops.model_backward(flat, x, saved, g.contiguous(), topo.rowptr, topo.col, topo.rowptr_t, topo.col_t,
                                    num_blocks, nc, topo.N, B, deterministic, plan)
        outs, off = [], 0
        for shp in shapes:
            n = 1
            for s in shp:
                n *= s
            outs.append(grads[off:off + n].view(shp))
            off += n
        return (None, None, None, None, None, None, None, None, None, *outs)


def gatres_model(x: Tensor, flat: Tensor, params: List[Tensor], topo, B: int, num_blocks: int, nc: int,
                 poison: Optional[Tensor] = None, deterministic: bool = False) -> Tensor:
    shapes = tuple(tuple(p.shape) for p in params)
    return _ModelFn.apply(x, flat.detach(), topo, B, num_blocks, nc, poison, shapes, deterministic, *params)
