"""Model registry for the GATRes configurations (drop-in for the GATRes rows of
/root/reference/gnn_pressure_estimation/ConfigModels.py:22-42, :123-178).

``select_model(args, name, reset_model_path)`` keeps the reference's contract:
it fills ``args.criterion / norm_type / use_data_edge_attrs / model_path`` and
returns ``(args, model)``.  The GATRes family and the plain ``gat`` baseline (same kernels) are built;
asking for one of the reference's other baseline models raises.
"""
from __future__ import annotations

import argparse
from typing import Callable, Dict, Optional, Tuple

import torch

from .GraphModels import GAT, GATResMeanConv

# (default variant name, num_blocks, nc, the authors' default checkpoint path)
_GATRES_VARIANTS: Dict[str, Tuple[str, int, int, str]] = {
    "gatres_small": ("GATResMeanConv_small_znorm_15b_32c", 15, 32,
                     r"experiments_logs\simple_test\gatres_znorm\best_GATResMeanConv_znorm_20235922.pth"),
    "gatres_large": ("GATRes_Large_znorm_25b_128c", 25, 128,
                     r"experiments_logs\simple_test\GATResMeanConvLarge_znorm_25b_128c_20235401_20230601_1754"
                     r"\best_GATResMeanConvLarge_znorm_25b_128c_20235401.pth"),
    "gatres_small_tough": ("GATResMeanConv_small_tough_znorm_15b_32c", 15, 32,
                           r"experiments_logs\simple_test\GATRes_small_tough_znorm_15b_32c"
                           r"\best_GATRes_small_tough_znorm_15b_32c_20233629.pth"),
}
_NOT_ON_HOT_PATH = ("gin", "graphconvwat", "chebnet", "mgcn", "gcn2")


def _configure(variant: str) -> Callable[[argparse.Namespace, Optional[str]], Tuple[argparse.Namespace, torch.nn.Module]]:
    default_name, num_blocks, nc, ckpt = _GATRES_VARIANTS[variant]

    def config(args: argparse.Namespace, test_model_variant_name: Optional[str] = None):
        args.model_path = ckpt
        args.criterion = "mse"
        args.use_data_edge_attrs = None
        args.norm_type = "znorm"
        return args, GATResMeanConv(name=test_model_variant_name or default_name, num_blocks=num_blocks, nc=nc)

    config.__name__ = f"config_{variant}"
    return config


config_gatres_small = _configure("gatres_small")
config_gatres_large = _configure("gatres_large")
config_gatres_small_tough = _configure("gatres_small_tough")


def config_gat(args: argparse.Namespace, test_model_variant_name: Optional[str] = None):
    """ConfigModels.py:96-103: the plain GAT baseline (shares the GATRes kernels)"""
    args.model_path = r"experiments_logs\simple_test\GAT\best_GAT_10b_32c_2h_20231827.pth"
    args.criterion = "mse"
    args.use_data_edge_attrs = None
    args.norm_type = "znorm"
    return args, GAT(name="GAT_10b_32c_2h" if test_model_variant_name is None else test_model_variant_name,
                     num_blocks=10, nc=32, in_channels=1, out_channels=1)


def select_model(args: argparse.Namespace, test_model_variant_name: Optional[str] = None,
                 reset_model_path: bool = False) -> Tuple[argparse.Namespace, torch.nn.Module]:
    """``args.model`` (default ``gatres_small``) -> (args with the model's defaults, model)."""
    which = getattr(args, "model", "gatres_small")
    previous_path = getattr(args, "model_path", None)
    if which in _NOT_ON_HOT_PATH:
        raise NotImplementedError(f"model {which!r} is a baseline of the reference and is not part of the "
                                  "B200 GATRes hot path; use gatres_small / gatres_large")
    # the reference's select_model (ConfigModels.py:133-178) cannot reach config_gatres_small_tough (:123-130);
    # it is selectable here because the training CLI offers it
    table = {"gatres_small": config_gatres_small, "gatres_large": config_gatres_large, "gat": config_gat,
             "gatres_small_tough": config_gatres_small_tough}
    if which not in table:
        raise NotImplementedError(f"Unknown model! Got {which}!")
    args, model = table[which](args, test_model_variant_name)
    if reset_model_path:
        args.model_path = previous_path
    return args, model
