"""`from layers.graph_loss_layer import ...` as in the reference's scripts (src/python/layers/graph_loss_layer.py)."""
from .loss_layers import GraphLossFunction, GraphLossLayer  # noqa: F401
from .. import pyDeform as _pd


def Finalize(src_V, src_F, src_E, src_to_graph, graph_V, rigidity, param_id):
    """graph_loss_layer.py:63-66 -- the sparse post-process of src/lib/linear.cc between the two normalisations."""
    pid = int(param_id.item()) if hasattr(param_id, "item") else int(param_id)
    _pd.NormalizeByTemplate(src_V, pid)
    _pd.SolveLinear(src_V, src_F, src_E, src_to_graph, graph_V, rigidity, 0)
    _pd.DenormalizeByTemplate(src_V, pid)
