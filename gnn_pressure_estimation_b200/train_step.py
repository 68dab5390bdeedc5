"""One training step of the reference loop as a fixed kernel sequence.

Mirrors the per-batch body of ``train_one_epoch``
(/root/reference/gnn_pressure_estimation/train.py:159-190): zero the masked
inputs, forward, MSE over the masked nodes, backward, Adam(lr 5e-4, L2 6e-6,
train.py:348,550-551) — everything on the device, on one stream, with static
buffers so the whole step is captured once into a CUDA graph and replayed.
The data-parallel variant inserts one gradient all-reduce (NCCL) between
backward and Adam.

Per step the host does: three H2D copies into the static input buffers (x, y,
mask), one graph launch, and (optionally) one 4-byte D2H read of the loss.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib, dp as _dp, metrics as _metrics, ops as _gops
from ._lib import ModelDesc, call, ptr, stream
from .GraphModels import GATResMeanConv
from .graph import Topology


class TrainStep:
    def __init__(self, model: GATResMeanConv, topo: Topology, batch: int, mask_count_per_snapshot: int,
                 lr: float = 5e-4, weight_decay: float = 6e-6, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, use_graph: bool = True, deterministic: bool = False,
                 grad_buckets: Optional[int] = None, peer_allreduce: Optional[bool] = None,
                 device_mask_seed: Optional[int] = None,
                 metrics: Optional["_metrics.MaskedMetrics"] = None,
                 share_state_with: Optional["TrainStep"] = None):
        if not topo.shares_one_csr:
            raise NotImplementedError("TrainStep runs the fused one-CSR stack; a template with self loops needs the "
                                      "module-by-module path (call the model eagerly: GATResMeanConv.forward)")
        self.model, self.topo, self.B = model, topo, int(batch)
        self.N, self.nc, self.nb = topo.N, model.nc, model.num_blocks
        self.M = self.B * self.N
        self.count = self.B * int(mask_count_per_snapshot)        # masked nodes per local batch
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        # device_mask_seed: draw the per-snapshot exact-count mask on the device every step (keyed by the seed and
        # the Adam step counter) instead of receiving a host mask; metrics: the seven reference metrics per step
        self.mask_count_per_snapshot = int(mask_count_per_snapshot)
        self.device_mask_seed = device_mask_seed
        self.metrics = metrics
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.use_graph = use_graph
        # gradient buckets for the data-parallel all-reduce: backward walks blocks nb-1 .. 0, and a bucket's
        # slice of the flat gradient buffer is reduced on a side stream while the next range still computes.
        # (deterministic mode finishes every gradient in one final reduction, so it has a single bucket.)
        if grad_buckets is None:
            grad_buckets = 3 if (self.world > 1 and not deterministic) else 1
        self.block_ranges = _dp.bucket_ranges(self.nb, 1 if deterministic else grad_buckets)
        self._comm_stream: Optional[torch.cuda.Stream] = None
        dev = topo.rowptr.device
        self.device = dev
        self.flat = model.flat_parameters()
        _dp.broadcast_parameters_(self.flat, self.pg)           # all replicas start from rank 0's weights
        P = self.flat.numel()
        self.P = P
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.zeros(self.M, **f32)            # static inputs (H2D targets)
        self.y = torch.zeros(self.M, **f32)
        self.mask = torch.zeros(self.M, dtype=torch.uint8, device=dev)
        self.xm = torch.empty(self.M, **f32)
        self.out = torch.empty(self.M, **f32)
        self.d_out = torch.empty(self.M, **f32)
        self.loss = torch.zeros(1, **f32)
        self._loss_part = torch.empty(1024, **f32)
        self.grads = torch.zeros(P, **f32)
        # data parallel, atomic gradient mode: fuse the gradient all-reduce into the Adam kernel over NVLink peer
        # memory (dp.PeerGradients); peer_allreduce=None tries it and falls back to the overlapped NCCL all-reduce
        self.peer: Optional[_dp.PeerGradients] = None
        self._parity = 0
        if self.world > 1 and not deterministic and peer_allreduce is not False and dev.type == "cuda":
            try:
                self.peer = _dp.PeerGradients(P, self.pg, dev)
            except Exception as e:                                 # symmetric memory unavailable on this system
                if peer_allreduce:
                    raise
                import warnings
                warnings.warn(f"peer-memory gradient all-reduce unavailable ({e!r}); using NCCL all-reduce")
        if share_state_with is not None:
            # a second step object for another batch size (the smaller last batch of an epoch, train.py:302 has no
            # drop_last) that continues the same optimisation: same parameters, same Adam moments, same step counter
            if share_state_with.model is not model:
                raise ValueError("share_state_with must wrap the same model")
            self.exp_avg, self.exp_avg_sq = share_state_with.exp_avg, share_state_with.exp_avg_sq
            self.step_count = share_state_with.step_count
        else:
            self.exp_avg = torch.zeros(P, **f32)
            self.exp_avg_sq = torch.zeros(P, **f32)
            self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.deterministic = deterministic
        self._plan = _gops.plan_tensors(topo)                  # keeps the plan tensors alive behind the raw pointers
        self.desc = _gops._desc(self.nb, self.nc, self.N, self.B, topo.rowptr, topo.col, topo.rowptr_t, topo.col_t, None,
                                deterministic, self._plan)
        lib = _lib.load()
        self.saved = torch.empty(int(lib.gatres_saved_floats(C.byref(self.desc))), **f32)
        self.scratch = torch.empty(int(lib.gatres_scratch_floats(C.byref(self.desc), 1)), **f32)
        self.partial = torch.empty(self.desc.slots * _gops.a4(P), **f32) if deterministic else None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._graphs = [None, None]                       # peer mode: one captured step per gradient-buffer parity
        # our kernel launches per step (mask, forward, loss, backward, Adam); measured in capture()/first run
        self.kernels_per_step = 0

    # ------------------------------------------------------------------ pieces
    def _enqueue(self) -> None:
        lib = _lib.load()
        n0 = lib.gatres_launch_count()
        self._enqueue_impl()
        self.kernels_per_step = int(lib.gatres_launch_count() - n0)

    def _enqueue_impl(self) -> None:
        s = stream()
        d = C.byref(self.desc)
        grads = self.grads if self.peer is None else self.peer.grads[self._parity]
        if self.device_mask_seed is not None:
            # rank-distinct streams: every rank masks its own shard independently (train.py:172 draws per batch)
            rank = torch.distributed.get_rank(self.pg) if self.pg is not None else 0
            call("gatres_generate_mask", (self.device_mask_seed + 0x9E3779B9 * rank) & (2 ** 64 - 1), 0,
                 ptr(self.step_count), None, self.B, self.N, self.mask_count_per_snapshot, ptr(self.mask), s)
        call("gatres_apply_mask", ptr(self.x), ptr(self.mask), ptr(self.xm), self.M, s)
        call("gatres_forward", d, ptr(self.flat), ptr(self.xm), ptr(self.out), ptr(self.saved), ptr(self.scratch), s)
        call("gatres_masked_mse", ptr(self.out), ptr(self.y), ptr(self.mask), self.M, self.count, ptr(self.d_out),
             ptr(self.loss), ptr(self._loss_part), s)
        if self.metrics is not None:
            self.metrics.update(self.out, self.y, self.mask)                # train.py:191-198, on the device
        # equal shard sizes and equal masked counts per snapshot -> mean of local means == global mean, so the
        # collective is a plain SUM (the 1/world factor is folded into the Adam kernel)
        dp = self.pg is not None and self.world > 1
        if not dp or len(self.block_ranges) == 1 or self.peer is not None:
            call("gatres_backward", d, ptr(self.flat), ptr(self.xm), ptr(self.saved), ptr(self.d_out),
                 ptr(self.partial), ptr(grads), ptr(self.scratch), s)
            if dp and self.peer is None:
                torch.distributed.all_reduce(self.grads, group=self.pg)
        else:
            self._backward_overlapped(d, s)
        if self.peer is not None:
            # all-reduce fused into Adam: every rank sums all ranks' gradient buffers through NVLink peer pointers
            call("gatres_adam_step_peer", ptr(self.flat), self.peer.grad_tables[self._parity], self.peer.flag_table,
                 self.peer.rank, self.world, ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.step_count),
                 ptr(self.peer.epoch), self.P, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                 1.0 / self.world, s)
            self._parity ^= 1
            return
        call("gatres_adam_step", ptr(self.flat), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.step_count), self.P, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
             1.0 / self.world, s)

    def _backward_overlapped(self, d, s) -> None:
        """Backward in block ranges; each range's finished slice of the flat gradient buffer is all-reduced
        on a communication stream while the compute stream continues with the next range.  Only the last
        bucket (block 0 + lin0) is exposed.  Works eagerly and under CUDA-graph capture (fork/join by events)."""
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        comm, compute = self._comm_stream, torch.cuda.current_stream()
        lib = _lib.load()
        for (k_hi, k_lo) in self.block_ranges:
            call("gatres_backward_range", d, ptr(self.flat), ptr(self.xm), ptr(self.saved), ptr(self.d_out),
                 ptr(self.partial), ptr(self.grads), ptr(self.scratch), k_hi, k_lo, s)
            lo, hi = _dp.bucket_slice(self.nb, self.P, k_hi, k_lo,
                                      lambda k: int(lib.gatres_param_offset_of_block(self.nb, self.nc, k)))
            comm.wait_stream(compute)
            with torch.cuda.stream(comm):
                torch.distributed.all_reduce(self.grads[lo:hi], group=self.pg)
        compute.wait_stream(comm)

    def capture(self, warmup: int = 1) -> None:
        """Run the step eagerly on a side stream (lazy CUDA/NCCL initialisation must
        not happen inside a capture), restore the optimizer state it touched, then
        capture the step into a CUDA graph."""
        if not self.use_graph:
            return
        state = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_count)]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._enqueue()
        torch.cuda.current_stream().wait_stream(side)
        for t, s in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.step_count), state):
            t.copy_(s)
        torch.cuda.synchronize(self.device)
        if self.peer is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g
            return
        start = self._parity                              # the warm-up steps advanced it identically on every rank
        for k in range(2):
            par = start ^ k
            self._parity = par
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()                           # flips self._parity
            self._graphs[par] = g
        self._parity = start
        self.graph = self._graphs[start]

    # ------------------------------------------------------------------- steps
    def load_inputs(self, x: Tensor, y: Tensor, mask: Optional[Tensor] = None) -> None:
        """H2D (or D2D) copy of one batch into the static buffers; x/y are the
        UNMASKED z-normed pressures ([B*N] or [B*N,1]), mask is bool/uint8 [B*N]
        (omit it when the mask is drawn on the device)."""
        self.x.copy_(x.reshape(-1), non_blocking=True)
        if y is not x:
            self.y.copy_(y.reshape(-1), non_blocking=True)
        else:
            self.y.copy_(self.x, non_blocking=True)                        # x = y in the reference (auxil.py:96-97)
        if mask is None:
            if self.device_mask_seed is None:
                raise ValueError("no mask given and no device_mask_seed configured")
            return
        if self.device_mask_seed is not None:
            raise ValueError("a host mask was given but this TrainStep draws its masks on the device")
        m = mask.reshape(-1)
        self.mask.copy_(m.view(torch.uint8) if m.dtype == torch.bool else m, non_blocking=True)

    def run(self) -> Tensor:
        """Enqueue one optimizer step on the resident inputs; returns the device loss scalar."""
        if self.graph is not None and self.peer is not None:
            self._graphs[self._parity].replay()
            self._parity ^= 1
        elif self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        return self.loss

    def step(self, x: Tensor, y: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        self.load_inputs(x, y, mask)
        return self.run()
