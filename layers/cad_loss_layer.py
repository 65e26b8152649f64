"""src/python/layers/cad_loss_layer.py"""
from meshode_b200.layers.cad_loss_layer import *  # noqa: F401,F403
