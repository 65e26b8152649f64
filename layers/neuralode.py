"""src/python/layers/neuralode.py (the slower variant of the same flow field; same classes here)."""
from meshode_b200.layers.neuralode import NeuralODE, ODEFunc, odeint_rk4  # noqa: F401
