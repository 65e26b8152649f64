"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

Two independent axes:

* **shape pairs** are independent units: rank r owns the contiguous block
  ``shard_range(n_pairs, r, world)``; there is no data-path collective, only a final gather of
  the per-pair results (``gather_pair_vertices``).
* **z-slabs of one large grid**: every rank holds the whole (replicated) target, builds voxel
  slices ``slab_range(N, r, world)`` (mo_template_create_slab), then ONE all-gather assembles the
  full field on every rank (``allgather_slabs``), because a vertex's trilinear cell needs slices
  z and z+1 wherever it moves.

The collectives work on whatever device the tensors live on: NCCL over NVLink for CUDA tensors,
gloo for the CPU tests (tests/test_sharding_gloo.py, world_size 2).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block partition: item i -> rank floor(i*world/n)."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def slab_range(N, rank, world):
    """Voxel slices [z0, z1) of rank ``rank`` (slab heights differ by at most one)."""
    return shard_range(N, rank, world)


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allgather_slabs(local, N, group=None):
    """``local``: tensor [z1-z0, N, N] holding this rank's slices of an N^3 field (any dtype).
    Returns the assembled [N, N, N] field on every rank (one all_gather; ragged slabs are padded
    to the tallest slab and trimmed)."""
    rank, world = _world(group)
    z0, z1 = slab_range(N, rank, world)
    if tuple(local.shape) != (z1 - z0, N, N):
        raise ValueError("rank %d expects a slab of shape %s, got %s" % (rank, (z1 - z0, N, N), tuple(local.shape)))
    if world == 1:
        return local.contiguous()
    hmax = max(slab_range(N, r, world)[1] - slab_range(N, r, world)[0] for r in range(world))
    send = local.contiguous()
    if z1 - z0 < hmax:
        pad = torch.zeros((hmax - (z1 - z0), N, N), dtype=local.dtype, device=local.device)
        send = torch.cat([send, pad], dim=0)
    recv = torch.empty((world, hmax, N, N), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(world * hmax, N, N), send, group=group)
    if N % world == 0:
        return recv.view(N, N, N)
    parts = []
    for r in range(world):
        a, b = slab_range(N, r, world)
        parts.append(recv[r, : b - a])
    return torch.cat(parts, dim=0)


def gather_pair_vertices(local_V, n_pairs, group=None, dst=0):
    """Final gather of the pair-sharded run: ``local_V`` is this rank's list of [n_i,3] float32
    result tensors for pairs ``shard_range(n_pairs, rank, world)``; rank ``dst`` receives the list
    of all ``n_pairs`` results in pair order (other ranks get None)."""
    rank, world = _world(group)
    lo, hi = shard_range(n_pairs, rank, world)
    if len(local_V) != hi - lo:
        raise ValueError("rank %d owns %d pairs, got %d results" % (rank, hi - lo, len(local_V)))
    if world == 1:
        return list(local_V)
    if local_V:
        dev = local_V[0].device
    elif dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())   # a rank that owns no pair still joins the NCCL collective
    else:
        dev = torch.device("cpu")
    counts = torch.tensor([v.shape[0] for v in local_V], dtype=torch.int64, device=dev)
    per = max(shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0] for r in range(world))
    cnt_pad = torch.zeros(per, dtype=torch.int64, device=dev)
    cnt_pad[: counts.numel()] = counts
    all_cnt = torch.empty((world, per), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_cnt.view(-1), cnt_pad, group=group)
    vmax = int(all_cnt.max().item()) if all_cnt.numel() else 0
    buf = torch.zeros((per, vmax, 3), dtype=torch.float32, device=dev)
    for i, v in enumerate(local_V):
        buf[i, : v.shape[0]] = v
    recv = torch.empty((world, per, vmax, 3), dtype=torch.float32, device=dev) if rank == dst else None
    if dist.get_backend(group) == "nccl":
        allr = torch.empty((world, per, vmax, 3), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allr.view(world * per, vmax, 3), buf, group=group)
        recv = allr if rank == dst else None
    else:
        dist.gather(buf, list(recv.unbind(0)) if rank == dst else None, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = shard_range(n_pairs, r, world)
        for i in range(b - a):
            out.append(recv[r, i, : int(all_cnt[r, i])].clone())
    return out


def layer_groups(N, world, layer=4):
    """Cyclic sharding of an N^3 field over ``world`` ranks by z-tile layers of ``layer`` slices: rank r builds layers
    r, r + world, ...  Returns the number of groups (a group = ``world`` consecutive layers = ``layer*world`` slices,
    contiguous in memory with rank r's piece at offset ``layer*r``) or 0 when N is not a multiple of ``layer*world``."""
    span = layer * world
    return N // span if N % span == 0 else 0


def _check_collectively(ok, message, group=None):
    """Raises on EVERY rank if any rank saw a problem (a one-sided raise would leave the others in a collective)."""
    rank, world = _world(group)
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok), group=group)
        ok = all(flags)
    if not ok:
        raise ValueError(message)


def build_template_sharded(tarV, tarF, grid_resolution, group=None, mode="auto", timings=False):
    """InitializeDeformTemplate with the grid build sharded over the ranks of ``group``: every rank builds its share
    of the voxel slices on its own GPU, all-gathers assemble the fields, and every rank ends up with a complete
    template.  Returns param_id (and a dict of CUDA-event timings when ``timings``).

    ``mode``: "cyclic" -- rank r builds z-tile layers r, r+world, ... (4 slices each), so every share spans the whole z
    range and costs the same wherever the surface lies; one in-place all-gather per group of ``world`` layers and
    field.  "slab" -- contiguous z-slabs (slices [r*N/world, (r+1)*N/world)), one all-gather per field; slabs through
    the middle of a shape cost more than polar ones.  "auto" -- cyclic when N is a multiple of 4*world and the backend
    is NCCL, else slab."""
    from . import capi
    from . import pyDeform as pd
    rank, world = _world(group)
    N = int(grid_resolution)
    ev = (lambda: torch.cuda.Event(enable_timing=True)) if timings else None
    if world == 1:
        e0, e1 = (ev(), ev()) if timings else (None, None)
        if timings:
            e0.record()
        pid = pd.InitializeDeformTemplate(tarV, tarF, 0, N)
        if timings:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            return pid, {"mode": "single", "build_ms": ms, "gather_ms": 0.0, "total_ms": ms}
        return pid
    _check_collectively(world <= N, "more ranks (%d) than voxel slices (%d)" % (world, N), group)
    nccl = dist.get_backend(group) == "nccl"
    groups = layer_groups(N, world, capi.LAYER_SLICES)
    if mode == "auto":
        mode = "cyclic" if (groups > 0 and nccl) else "slab"
    _check_collectively(mode in ("cyclic", "slab") and (mode != "cyclic" or (groups > 0 and nccl)),
                        "cyclic sharding needs NCCL and grid_resolution %% (%d * world) == 0" % capi.LAYER_SLICES, group)
    V = tarV.cuda() if not tarV.is_cuda else tarV
    F = tarF.cuda() if not tarF.is_cuda else tarF
    s = torch.cuda.current_stream().cuda_stream
    e0, e1, e2 = (ev(), ev(), ev()) if timings else (None, None, None)
    if timings:
        e0.record()
    if mode == "cyclic":
        pid = capi.template_create_layers(V.data_ptr(), V.shape[0], F.data_ptr(), F.shape[0], N, rank, world, s)
        if timings:
            e1.record()
        span, L = capi.LAYER_SLICES * world, capi.LAYER_SLICES
        fields = pd.GridViews(pid)
        # all-gathers IN PLACE on the template's own fields; grouped so that NCCL issues them as one batch
        try:
            ctx = dist._coalescing_manager(group=group, device=V.device, async_ops=False)
        except (AttributeError, TypeError):
            ctx = None
        def issue():
            for field in fields:
                for j in range(groups):
                    chunk = field[j * span:(j + 1) * span]
                    dist.all_gather_into_tensor(chunk, chunk[rank * L:(rank + 1) * L], group=group)
        if ctx is not None:
            with ctx:
                issue()
        else:
            issue()
    else:
        z0, z1 = slab_range(N, rank, world)
        pid = capi.template_create_slab(V.data_ptr(), V.shape[0], F.data_ptr(), F.shape[0], N, z0, z1, s)
        if timings:
            e1.record()
        if N % world == 0 and nccl:
            # equal slabs: all-gather IN PLACE on the template's own fields (rank r's slab already sits at offset r)
            for field in pd.GridViews(pid):
                dist.all_gather_into_tensor(field, field[z0:z1], group=group)
        else:
            g64, g32, idx = pd.GetGrid(pid, z0, z1)
            full64 = allgather_slabs(g64[z0:z1], N, group)
            full32 = allgather_slabs(g32[z0:z1], N, group)
            fulli = allgather_slabs(idx[z0:z1], N, group)
            pd.SetGrid(pid, full64.contiguous(), full32.contiguous(), fulli.contiguous())
    if timings:
        e2.record()
        torch.cuda.synchronize()
        return pid, {"mode": mode, "build_ms": e0.elapsed_time(e1), "gather_ms": e1.elapsed_time(e2),
                     "total_ms": e0.elapsed_time(e2)}
    return pid
