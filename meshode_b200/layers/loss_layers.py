"""Device-native loss layers: the host-side mirror of the reference's src/python/layers.

Same class names, constructor/forward signatures and loss definitions as
  rigid_loss_layer.py:7-44, graph_loss_layer.py:9-66, graph_loss2_layer.py:9-77,
  cad_loss_layer.py:7-48, reverse_loss_layer.py:10-26
but the tensors stay on the GPU: there is no ``.data.cpu().numpy()`` round trip
(graph_loss_layer.py:15,36), and forward + backward of the distance and edge terms are ONE kernel
launch (mo_loss_forward_backward) whose gradient is kept for autograd's backward.

Differences from the reference, all additive or bug-for-bug documented:
  * ``grid_resolution`` keyword (the reference hard-codes 64: rigid_loss_layer.py:35);
  * tensors handed to a layer's constructor are moved to the layer's device; ``src_V`` is still
    normalised IN PLACE as in the reference when it already lives there;
  * CadLossFunction.forward passes ``param_id`` (the reference omits it, cad_loss_layer.py:10-11,
    and would raise TypeError);
  * the scalar loss is accumulated in float64 and returned as float32;
  * forward/backward use the edge connectivity captured by Store*Information: the ``src_F`` / ``src_E`` handed to
    ``forward`` (or to the ``*LossFunction.apply``) are accepted for the reference's signature and not re-read
    (the reference iterates the caller's tensors against the stored rest vectors: rigid_layer.cc:113-130).
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import pyDeform

GRAPH_MASK = 0.5 * 0.03 * 0.03   # graph_loss_layer.py:18


def _pid(p):
    return int(p.item()) if isinstance(p, torch.Tensor) else int(p)


class _FusedLoss(Function):
    """loss = 0.5*sum(dist^2) + w_edge*0.5*sum(edge^2); d loss/dV = mask*dist_bwd + w_edge*edge_bwd."""

    @staticmethod
    def forward(ctx, V, dist_pid, edge_pid, w_edge, mask_threshold):
        Vc = V.detach().contiguous()
        need = V.requires_grad
        loss, grad = pyDeform.LossForwardBackward(Vc, dist_pid, edge_pid, w_edge, mask_threshold, True, need)
        if need:
            ctx.save_for_backward(grad)
        return loss.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_h):
        (grad,) = ctx.saved_tensors
        return grad_h * grad, None, None, None, None


def _function_forward(ctx, V, dist_pid, edge_pid, w_edge, mask_threshold):
    """forward of the exported autograd Functions: one fused launch gives the loss and its gradient, the gradient is
    kept for backward (the reference recomputes it there: rigid_loss_layer.py:20-27)."""
    loss, grad = pyDeform.LossForwardBackward(V.detach().contiguous(), dist_pid, edge_pid, w_edge, mask_threshold, True, True)
    ctx.save_for_backward(grad)
    return loss.to(torch.float32)


def _function_backward(ctx, grad_h, n_inputs):
    (grad,) = ctx.saved_tensors
    return (grad_h * grad,) + (None,) * (n_inputs - 1)


def _to_device(t, device):
    return t if t.device == device else t.to(device)


class _TemplateLayer(nn.Module):
    def __init__(self, device=None):
        super().__init__()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("meshode_b200 layers run on a CUDA device (no CPU fallback)")

    def _normalized_copy(self, V, pid):
        """NormalizeByTemplate in place (as the reference does to its argument) and a device copy."""
        pyDeform.NormalizeByTemplate(V, pid)
        return _to_device(V.detach(), self.device).contiguous()


# ---- rigid_loss_layer.py --------------------------------------------------------------------------
class RigidLossFunction(Function):
    """rigid_loss_layer.py:7-27: ``RigidLossFunction.apply(src_V, src_F, param_id)``, differentiable in src_V.  The edge
    connectivity is the one captured by StoreRigidityInformation (``src_F`` is accepted for the signature)."""

    @staticmethod
    def forward(ctx, src_V, src_F, param_id):
        pid = _pid(param_id)
        return _function_forward(ctx, src_V, pid, pid, 1.0, 0.0)

    @staticmethod
    def backward(ctx, grad_h):
        return _function_backward(ctx, grad_h, 3)


class RigidLossLayer(_TemplateLayer):
    def __init__(self, src_V, src_F, tar_V, tar_F, grid_resolution=64, device=None):
        super().__init__(device)
        with torch.cuda.device(self.device):
            pid = pyDeform.InitializeDeformTemplate(_to_device(tar_V, self.device), _to_device(tar_F, self.device), 0,
                                                    grid_resolution)
            self.param_id = torch.tensor(pid)
            Vn = self._normalized_copy(src_V, pid)
            pyDeform.StoreRigidityInformation(Vn, _to_device(src_F, self.device), pid)

    def forward(self, src_V, src_F=None):
        pid = _pid(self.param_id)
        return _FusedLoss.apply(src_V, pid, pid, 1.0, 0.0)


def Finalize(src_V, param_id):
    """rigid_loss_layer.py:43-44 / cad_loss_layer.py:47-48."""
    pyDeform.DenormalizeByTemplate(src_V.data if isinstance(src_V, nn.Parameter) else src_V, _pid(param_id))


# ---- graph_loss_layer.py --------------------------------------------------------------------------
class GraphLossFunction(Function):
    """graph_loss_layer.py:9-43 (distance gradient masked to vertices with 0.5*d^2 < 0.5*0.03^2)."""

    @staticmethod
    def forward(ctx, src_V, src_E, rigidity2, param_id):
        pid = _pid(param_id)
        return _function_forward(ctx, src_V, pid, pid, float(rigidity2), GRAPH_MASK)

    @staticmethod
    def backward(ctx, grad_h):
        return _function_backward(ctx, grad_h, 4)


class GraphLossLayer(_TemplateLayer):
    def __init__(self, src_V, src_E, tar_V, tar_F, rigidity, d=None, grid_resolution=64):
        super().__init__(d if d is not None and torch.device(d).type == "cuda" else None)
        with torch.cuda.device(self.device):
            pid = pyDeform.InitializeDeformTemplate(_to_device(tar_V, self.device), _to_device(tar_F, self.device), 0,
                                                    grid_resolution)
            self.param_id = torch.tensor(pid)
            Vn = self._normalized_copy(src_V, pid)
            pyDeform.StoreGraphInformation(Vn, _to_device(src_E, self.device), pid)
        self.rigidity2 = torch.tensor(float(rigidity) * float(rigidity))

    def forward(self, src_V, src_E=None):
        pid = _pid(self.param_id)
        return _FusedLoss.apply(src_V, pid, pid, float(self.rigidity2), GRAPH_MASK)


# ---- graph_loss2_layer.py -------------------------------------------------------------------------
class GraphLoss2Function(Function):
    """graph_loss2_layer.py:9-41."""

    @staticmethod
    def forward(ctx, V1, E1, rigidity2, param_id1, param_id2):
        # distance to the OTHER mesh (param_id2), edges of its own (param_id1): graph_loss2_layer.py:18-19
        return _function_forward(ctx, V1, _pid(param_id2), _pid(param_id1), float(rigidity2), 0.0)

    @staticmethod
    def backward(ctx, grad_h):
        return _function_backward(ctx, grad_h, 5)


class GraphLoss2Layer(_TemplateLayer):
    def __init__(self, V1, F1, graph_V1, graph_E1, V2, F2, graph_V2, graph_E2, rigidity, d=None, grid_resolution=64):
        super().__init__(d if d is not None and torch.device(d).type == "cuda" else None)
        dev = self.device
        with torch.cuda.device(dev):
            p1 = pyDeform.InitializeDeformTemplate(_to_device(V1, dev), _to_device(F1, dev), 0, grid_resolution)
            p2 = pyDeform.InitializeDeformTemplate(_to_device(V2, dev), _to_device(F2, dev), 0, grid_resolution)
            self.param_id1, self.param_id2 = torch.tensor(p1), torch.tensor(p2)
            g1 = self._normalized_copy(graph_V1, p1)
            g2 = self._normalized_copy(graph_V2, p2)
            pyDeform.StoreGraphInformation(g1, _to_device(graph_E1, dev), p1)
            pyDeform.StoreGraphInformation(g2, _to_device(graph_E2, dev), p2)
        self.rigidity2 = torch.tensor(float(rigidity) * float(rigidity))

    def forward(self, V1, E1, V2, E2, direction):
        if direction == 0:
            return _FusedLoss.apply(V1, _pid(self.param_id2), _pid(self.param_id1), float(self.rigidity2), 0.0)
        return _FusedLoss.apply(V2, _pid(self.param_id1), _pid(self.param_id2), float(self.rigidity2), 0.0)


# ---- cad_loss_layer.py ----------------------------------------------------------------------------
class CadLossFunction(Function):
    """cad_loss_layer.py:7-27 (with the param_id the reference forgets to pass to its forward calls, :10-11)."""

    @staticmethod
    def forward(ctx, src_V, src_F, src_E, param_id):
        pid = _pid(param_id)
        return _function_forward(ctx, src_V, pid, pid, 1.0, 0.0)

    @staticmethod
    def backward(ctx, grad_h):
        return _function_backward(ctx, grad_h, 4)


class CadLossLayer(_TemplateLayer):
    def __init__(self, src_V, src_F, src_E, tar_V, tar_F, grid_resolution=64, device=None):
        super().__init__(device)
        with torch.cuda.device(self.device):
            pid = pyDeform.InitializeDeformTemplate(_to_device(tar_V, self.device), _to_device(tar_F, self.device), 0,
                                                    grid_resolution)
            self.param_id = torch.tensor(pid)
            Vn = self._normalized_copy(src_V, pid)
            pyDeform.StoreCadInformation(Vn, _to_device(src_F, self.device), _to_device(src_E, self.device), pid)

    def forward(self, src_V, src_F=None, src_E=None):
        pid = _pid(self.param_id)
        return _FusedLoss.apply(src_V, pid, pid, 1.0, 0.0)


# ---- reverse_loss_layer.py ------------------------------------------------------------------------
class ReverseLossLayer(nn.Module):
    """loss = 0.5 * sum |src_V[nn(tar)] - tar_V|^2 with nn = nearest SOURCE vertex of every target
    vertex (reverse_loss_layer.py:15-24: cKDTree(src).query(tar)); differentiable through the gather.
    The exact nearest-vertex search runs on the GPU (mo_nearest_vertex, FP64 like cKDTree)."""

    def forward(self, src_V, tar_V, device=None):
        tar = _to_device(tar_V.detach(), src_V.device).contiguous()
        ii = pyDeform.NearestVertex(tar, src_V.detach().contiguous())
        diff = src_V[ii.long()] - tar
        return 0.5 * (diff * diff).sum()
