"""src/python/layers/reverse_loss_layer.py"""
from meshode_b200.layers.reverse_loss_layer import *  # noqa: F401,F403
from meshode_b200.layers.reverse_loss_layer import ReverseLossLayer  # noqa: F401
