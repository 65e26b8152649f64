"""`from layers.reverse_loss_layer import ...` as in the reference's scripts (src/python/layers/reverse_loss_layer.py)."""
from .loss_layers import ReverseLossLayer  # noqa: F401
