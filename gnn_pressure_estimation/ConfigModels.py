"""``gnn_pressure_estimation.ConfigModels`` -> B200 implementation."""
from gnn_pressure_estimation_b200.ConfigModels import (  # noqa: F401
    config_gat, config_gatres_large, config_gatres_small, config_gatres_small_tough, select_model)
