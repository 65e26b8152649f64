"""Whole-pair deformation on the GPU: template build + fused Adam loop, batched over pairs,
sharded over ranks.  This is the device-native form of src/python/rigid_deform.py:25-44.

The loops run inside libmeshode_b200.so (mo_deform_batch_adam: one persistent CTA per pair);
this module only marshals torch tensors and partitions pair indices across processes.
"""
import ctypes as C

import torch

from . import capi
from . import pyDeform as pd


def _stream():
    return torch.cuda.current_stream().cuda_stream


_SCHEDULES = {"auto": 0, "cta": capi.DEFORM_CTA_ONLY, "cluster": capi.DEFORM_CLUSTER_ONLY}


def deform_batch_adam(V_list, dist_pids, edge_pids, iters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, exact=True,
                      schedule="auto"):
    """In-place Adam optimisation of each normalised source ``V_list[i]`` (CUDA float32 [n_i,3])
    against template ``dist_pids[i]`` with the edges stored in template ``edge_pids[i]``.
    ``exact=True`` (default) reproduces the float32 CPU loop bit for bit; ``exact=False`` sums the edge
    term over distinct neighbours (same value to ~1e-10 per term, 1.4x faster), which Adam's chaotic
    sensitivity turns into ~2e-4 Chamfer after 10 000 iterations -- the same as a 1-ulp change of the input.
    ``schedule`` (exact loop): "auto" runs full waves one CTA per pair and a partial wave on thread-block clusters
    (several SMs per pair); "cta" / "cluster" pin one of the two kernels.  Same bits either way."""
    B = len(V_list)
    assert len(dist_pids) == B and len(edge_pids) == B
    for v in V_list:
        if not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous() and v.dim() == 2 and v.shape[1] == 3):
            raise ValueError("V must be contiguous CUDA float32 [n,3]")
    dp = (C.c_int * B)(*[int(p) for p in dist_pids])
    ep = (C.c_int * B)(*[int(p) for p in edge_pids])
    vp = (C.c_void_p * B)(*[v.data_ptr() for v in V_list])
    capi.check(capi.lib().mo_deform_batch_adam(dp, ep, vp, B, int(iters), float(lr), float(betas[0]), float(betas[1]),
                                               float(eps), (capi.DEFORM_EXACT | _SCHEDULES[schedule]) if exact else 0, _stream()))


def deform_adam_large(V, dist_pid, edge_pid, iters, lr=1e-3, w_edge=1.0, mask_threshold=0.0, betas=(0.9, 0.999),
                      eps=1e-8):
    """Same loop for one pair of any size (two launches per iteration)."""
    capi.check(capi.lib().mo_deform_adam_large(int(dist_pid), int(edge_pid), V.data_ptr(), V.shape[0], float(w_edge),
                                               float(mask_threshold), int(iters), float(lr), float(betas[0]),
                                               float(betas[1]), float(eps), _stream()))


class PairBatch:
    """A batch of (source, target) pairs resident on one GPU: templates built, sources normalised,
    rest edges stored -- what RigidLossLayer.__init__ does (rigid_loss_layer.py:31-38), per pair."""

    _side_streams = {}

    def __init__(self, pairs, grid_resolution=64, device=None, n_streams=4):
        """``pairs``: iterable of (srcV, srcF, tarV, tarF) torch tensors (CPU pinned or CUDA).
        The per-pair set-up kernels are spread over ``n_streams`` side streams so that one pair's
        small binning kernels overlap another pair's tile kernel; the caller's current stream
        waits for all of them before this constructor returns."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N = int(grid_resolution)
        self.V, self.F, self.pids = [], [], []
        self._keep = []
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            key = (self.device.index, n_streams)
            if key not in PairBatch._side_streams:
                PairBatch._side_streams[key] = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
            sides = PairBatch._side_streams[key] if n_streams > 1 else [main]
            for s in sides:
                s.wait_stream(main)
            for i, (srcV, srcF, tarV, tarF) in enumerate(pairs):
                with torch.cuda.stream(sides[i % len(sides)]):
                    sV = srcV.to(self.device, non_blocking=True).contiguous()
                    sF = srcF.to(self.device, non_blocking=True).contiguous()
                    tV = tarV.to(self.device, non_blocking=True).contiguous()
                    tF = tarF.to(self.device, non_blocking=True).contiguous()
                    if sV.data_ptr() == srcV.data_ptr():
                        sV = sV.clone()
                    pid = pd.InitializeDeformTemplate(tV, tF, 0, self.N)
                    pd.NormalizeByTemplate(sV, pid)
                    pd.StoreRigidityInformation(sV, sF, pid)
                self.V.append(sV); self.F.append(sF); self.pids.append(pid)
                self._keep.append((tV, tF))
            for s in sides:
                main.wait_stream(s)
            for t in self.V + self.F + [x for kv in self._keep for x in kv]:
                t.record_stream(main)
            self._keep = []

    def deform(self, iters=10000, lr=1e-3, exact=True, schedule="auto"):
        with torch.cuda.device(self.device):
            small = [i for i, v in enumerate(self.V) if v.shape[0] <= 6144]
            large = [i for i, v in enumerate(self.V) if v.shape[0] > 6144]
            if small:
                deform_batch_adam([self.V[i] for i in small], [self.pids[i] for i in small], [self.pids[i] for i in small],
                                  iters, lr, exact=exact, schedule=schedule)
            for i in large:
                deform_adam_large(self.V[i], self.pids[i], self.pids[i], iters, lr)

    def finalize(self):
        """Finalize (rigid_loss_layer.py:43-44): denormalise in place; returns the vertex tensors."""
        with torch.cuda.device(self.device):
            for v, pid in zip(self.V, self.pids):
                pd.DenormalizeByTemplate(v, pid)
        return self.V

    def release(self):
        """Returns the templates' buffers to the pool in the order of the current stream (the one deform() and
        finalize() were enqueued on)."""
        with torch.cuda.device(self.device):
            for pid in self.pids:
                pd.DestroyTemplate(pid)
        self.pids = []


from .sharding import shard_range  # noqa: E402,F401  (kept here for callers of engine.shard_range)
