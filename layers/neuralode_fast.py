"""src/python/layers/neuralode_fast.py -- also the module path under which NeuralODE / ODEFunc objects are pickled
(torch.save({'func': func, ...}) in cad_neural_deform2.py:108), so checkpoints move between the two code bases."""
from meshode_b200.layers.neuralode import NeuralODE, ODEFunc, odeint_rk4  # noqa: F401
