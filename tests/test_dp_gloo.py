"""world_size-2 gloo tests of the data-parallel host logic on CPU (the per-rank
compute is the oracle here — the CUDA path has no CPU mode; the same contract is
checked on the GPU in test_gpu_parity.py::test_gradient_is_mean_over_shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_pressure_estimation_b200 import dp
from helpers import load_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors])


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from oracle import gatres_oracle as O
    r, w, _ = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    c = load_case("ctown_small_15b_32c_B8")
    N, B = c["N"], c["B"]
    # rank 1 starts from different weights; broadcast must repair that
    model = O.make_oracle(3, 32, seed=0 if rank == 0 else 7)
    params = list(model.parameters())
    flat = _flat([p.detach() for p in params])
    assert dp.replicas_in_sync(flat) is False
    dp.broadcast_parameters_(flat)
    off = 0
    with torch.no_grad():
        for p in params:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    assert dp.replicas_in_sync(flat)

    lo, hi = dp.shard_bounds(B, rank, world)
    sl = slice(lo * N, hi * N)
    eib = O.collate_edge_index(c["edge_index"], N, hi - lo)
    _, loss, grads = O.train_step_loss_and_grads(model, c["x"][sl], c["y"][sl], c["mask"][sl], eib)
    g = _flat([grads[k] for k, _ in model.named_parameters()])
    dp.allreduce_gradients_(g)

    # every rank applies the same Adam step -> replicas stay bit-identical
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    O.adam_reference_step(flat, g, m, v, step=1)
    assert dp.replicas_in_sync(flat)
    if rank == 0:
        ret["grad"] = g.clone()
        ret["loss"] = float(loss)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_step():
    from oracle import gatres_oracle as O
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        g_dp = ret["grad"]
    c = load_case("ctown_small_15b_32c_B8")
    model = O.make_oracle(3, 32, seed=0)
    eib = O.collate_edge_index(c["edge_index"], c["N"], c["B"])
    _, _, grads = O.train_step_loss_and_grads(model, c["x"], c["y"], c["mask"], eib)
    g_full = _flat([grads[k] for k, _ in model.named_parameters()])
    assert float((g_dp - g_full).abs().max()) <= 1e-5 * float(g_full.abs().max())


def test_shard_bounds():
    assert [dp.shard_bounds(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    assert dp.shard_bounds(16384, 3, 8) == (6144, 8192)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)


def test_uneven_shards_cover_the_set_once():
    """evaluation shards (a validation split need not divide by the world size; empty shards are legal)"""
    for total, world in ((10, 4), (3, 8), (64, 8), (0, 2), (1001, 8)):
        b = [dp.shard_bounds_uneven(total, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("nb,buckets", [(15, 3), (15, 4), (25, 5), (2, 3), (1, 1), (0, 2), (15, 1)])
def test_gradient_buckets_partition_the_flat_buffer(nb, buckets):
    """Backward walks blocks nb-1..0; the slices that become final after each range must tile [0, P)
    from the top down (layout: lin0 | block 0 .. nb-1 | lin1), with no gap and no overlap."""
    nc = 32
    block = 4 * nc * nc + 9 * nc
    P = 2 * nc + nb * block + nc + 1
    off = lambda k: 2 * nc + min(max(k, 0), nb) * block if k >= 0 else 0   # same rule as gatres_param_offset_of_block
    ranges = dp.bucket_ranges(nb, buckets)
    assert len(ranges) == (max(1, min(buckets, nb)) if nb > 0 else 1)
    if nb > 0:
        assert ranges[0][0] == nb - 1 and ranges[-1][1] == 0
        assert all(a[1] == b[0] + 1 for a, b in zip(ranges, ranges[1:]))        # descending, contiguous
        sizes = [hi - lo + 1 for hi, lo in ranges]
        assert max(sizes) - min(sizes) <= 1
    top = P
    for k_hi, k_lo in ranges:
        lo, hi = dp.bucket_slice(nb, P, k_hi, k_lo, off)
        assert hi == top and lo < hi
        top = lo
    assert top == 0
