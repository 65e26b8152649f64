"""src/python/layers/graph_loss_layer.py"""
from meshode_b200.layers.graph_loss_layer import *  # noqa: F401,F403
from meshode_b200.layers.graph_loss_layer import Finalize, GraphLossFunction, GraphLossLayer  # noqa: F401
