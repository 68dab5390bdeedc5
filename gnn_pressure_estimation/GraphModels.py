"""``gnn_pressure_estimation.GraphModels`` -> B200 implementation (see gnn_pressure_estimation_b200/GraphModels.py)."""
from gnn_pressure_estimation_b200.GraphModels import *  # noqa: F401,F403
from gnn_pressure_estimation_b200.GraphModels import (  # noqa: F401
    GAT, GATConvNet, GATResMeanConv, GResBlockConv, GResBlockMeanConv)
