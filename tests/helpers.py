import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| (SURVEY §8c definition)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, rtol, what=""):
    a, b = a.detach().cpu(), b.detach().cpu()
    e = rel_err(a, b)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"
    assert torch.allclose(a.double(), b.double(), rtol=rtol, atol=rtol * float(b.abs().max() + 1e-30)), what


def random_directed_graph(n, e, seed, allow_self_loops=False):
    """asymmetric multigraph edge_index (int64 [2,e]); exercises the transposed CSR."""
    rng = np.random.RandomState(seed)
    src = rng.randint(0, n, size=e)
    dst = rng.randint(0, n, size=e)
    if not allow_self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    return torch.from_numpy(np.stack([src, dst]).astype(np.int64))
