"""Top-level ``layers`` package: the import path of the reference's src/python/layers, so that its scripts
(``from layers.rigid_loss_layer import RigidLossLayer, Finalize``; src/python/rigid_deform.py:10) run unmodified with the
repository root on PYTHONPATH.  Re-exports meshode_b200.layers (device-native)."""
