"""src/python/layers/graph_loss2_layer.py"""
from meshode_b200.layers.graph_loss2_layer import *  # noqa: F401,F403
from meshode_b200.layers.graph_loss2_layer import Finalize, GraphLoss2Function, GraphLoss2Layer  # noqa: F401
