"""Host-side mirror of the reference's src/python/layers package (device-native)."""
from .loss_layers import (CadLossFunction, CadLossLayer, Finalize, GraphLoss2Function, GraphLoss2Layer,  # noqa: F401
                          GraphLossFunction, GraphLossLayer, ReverseLossLayer, RigidLossFunction, RigidLossLayer)
from .neuralode import NeuralODE, ODEFunc  # noqa: F401
