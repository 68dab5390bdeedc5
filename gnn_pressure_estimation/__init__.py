"""Import shim: lets reference callers keep their import lines
(``from gnn_pressure_estimation.GraphModels import GATResMeanConv``,
/root/reference/gnn_pressure_estimation/ConfigModels.py:11-19, evaluation.py:16-19)
while getting the B200 implementation."""
