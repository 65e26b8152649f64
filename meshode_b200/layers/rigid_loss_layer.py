"""`from layers.rigid_loss_layer import ...` as in the reference's scripts (src/python/layers/rigid_loss_layer.py)."""
from .loss_layers import RigidLossFunction, RigidLossLayer, Finalize  # noqa: F401
