"""`from layers.cad_loss_layer import ...` as in the reference's scripts (src/python/layers/cad_loss_layer.py)."""
from .loss_layers import CadLossFunction, CadLossLayer, Finalize  # noqa: F401
