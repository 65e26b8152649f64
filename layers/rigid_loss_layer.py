"""src/python/layers/rigid_loss_layer.py"""
from meshode_b200.layers.rigid_loss_layer import *  # noqa: F401,F403
from meshode_b200.layers.rigid_loss_layer import Finalize, RigidLossFunction, RigidLossLayer  # noqa: F401
