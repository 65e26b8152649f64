"""NeuralODE flow field of the reference (src/python/layers/neuralode_fast.py:7-64) in stock PyTorch.

The reference integrates a 4->50->50->50->3 LeakyReLU MLP with ``torchdiffeq.odeint(method="rk4")``
over the two-point time grid [0,1] (or [1,0] for the inverse).  torchdiffeq is not installed here, so
the fixed-grid solver is restated: its "rk4" method takes ONE step per grid interval with the 3/8-rule
Runge-Kutta tableau (torchdiffeq/_impl/rk_common.py::rk4_alt_step_func -- recalled, unpinned: the
package is absent from the environment).  This module is host-side glue; only its loss runs on the
meshode_b200 hot path.
"""
import torch
from torch import nn


class ODEFunc(nn.Module):
    def __init__(self):
        super().__init__()
        m = 50
        self.net = nn.Sequential(nn.Linear(4, m), nn.LeakyReLU(), nn.Linear(m, m), nn.LeakyReLU(), nn.Linear(m, m),
                                 nn.LeakyReLU(), nn.Linear(m, 3))
        for mod in self.net.modules():
            if isinstance(mod, nn.Linear):
                nn.init.normal_(mod.weight, mean=0, std=1e-1)
                nn.init.constant_(mod.bias, val=0)

    def forward(self, t, y):
        yt = torch.cat((y, t.reshape(1, 1).expand(y.shape[0], 1)), 1)
        return self.net(yt - 0.5)


def rk4_38_step(func, t0, dt, y0):
    """One step of the 3/8-rule fourth-order Runge-Kutta method."""
    k1 = func(t0, y0)
    k2 = func(t0 + dt / 3, y0 + dt * k1 / 3)
    k3 = func(t0 + dt * 2 / 3, y0 + dt * (k2 - k1 / 3))
    k4 = func(t0 + dt, y0 + dt * (k1 - k2 + k3))
    return y0 + dt * (k1 + 3 * (k2 + k3) + k4) / 8


def odeint_rk4(func, y0, t):
    """Fixed-grid integration over the time points ``t``; returns the solution at every point."""
    ys = [y0]
    for i in range(t.shape[0] - 1):
        ys.append(rk4_38_step(func, t[i], t[i + 1] - t[i], ys[-1]))
    return torch.stack(ys)


class NeuralODE:
    def __init__(self, device=torch.device("cpu")):
        self.timing = torch.tensor([0.0, 1.0], dtype=torch.float32, device=device)
        self.timing_inv = torch.tensor([1.0, 0.0], dtype=torch.float32, device=device)
        self.func = ODEFunc().to(device)
        self.device = device

    def to_device(self, device):
        self.func = self.func.to(device)
        self.timing = self.timing.to(device)
        self.timing_inv = self.timing_inv.to(device)
        self.device = device

    def parameters(self):
        return self.func.parameters()

    def forward(self, u):
        return odeint_rk4(self.func, u, self.timing)[1]

    def inverse(self, u):
        return odeint_rk4(self.func, u, self.timing_inv)[1]

    def integrate(self, u, t1, t2, device):
        return odeint_rk4(self.func, u, torch.tensor([t1, t2], dtype=torch.float32, device=device))[1]


# Checkpoints: the reference pickles the objects themselves -- torch.save({'func': func, 'optim': optimizer})
# (src/python/cad_neural_deform2.py:108) -- and its consumer reads checkpoint['func'] / ['optim'] back as objects
# (src/python/cad_neural_animate.py:52-57).  Pickle records the defining module, so the classes carry the reference's
# module path (served by the top-level ``layers/neuralode_fast.py``): a checkpoint written here loads in the reference's
# tooling and the other way round.
ODEFunc.__module__ = "layers.neuralode_fast"
NeuralODE.__module__ = "layers.neuralode_fast"


def save_checkpoint(path, func, optimizer):
    """The reference's layout: {'func': NeuralODE object, 'optim': optimizer object}."""
    import layers.neuralode_fast  # noqa: F401  (the pickled module path must resolve to these classes)
    torch.save({"func": func, "optim": optimizer}, path)


def load_checkpoint(path, device, lr=1e-3):
    """Returns (func, optimizer) from a checkpoint in the reference's layout (pickled objects) or in the
    state_dict layout written by round 1 of this repository ({'func': ODEFunc.state_dict(), 'optim':
    Adam.state_dict()})."""
    import layers.neuralode_fast  # noqa: F401
    ck = torch.load(path, map_location=device, weights_only=False)
    f, o = ck["func"], ck["optim"]
    if isinstance(f, dict):
        func = NeuralODE(device)
        func.func.load_state_dict(f)
    else:
        func = f
        func.to_device(device)
    if isinstance(o, dict):
        optimizer = torch.optim.Adam(func.parameters(), lr=lr)
        optimizer.load_state_dict(o)
    else:
        optimizer = o
    return func, optimizer
